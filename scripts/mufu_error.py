"""Bias / rms / max error of the MUFU approximations used by the default fp32 math (bl_mufu_error), and the
per-component gradient error of the occu kernels near the mode at config-2 size (diagnostic for DESIGN.md)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
from biolith_b200 import _lib

lib = _lib.load()
for which, name, ranges in ((0, "ex2.approx rel", [(-24, 0), (0, 24), (-1, 0), (0, 1), (-8, 8)]),
                            (1, "lg2.approx abs", [(1, 2), (1, 256), (2, 1e6)]),
                            (2, "rcp+newton rel", [(1, 2), (1, 256), (1, 1e20)])):
    for lo, hi in ranges:
        m, r, x = C.c_double(), C.c_double(), C.c_double()
        _lib.check(lib.bl_mufu_error(0, which, lo, hi, 1 << 26, C.byref(m), C.byref(r), C.byref(x)), "bl_mufu_error")
        print(f"{name:16s} x in [{lo:g}, {hi:g}]: mean {m.value:+.3e}  rms {r.value:.3e}  max {x.value:.3e}")

if "--grad" in sys.argv:
    from oracle import c_oracle
    S = 1_000_000
    data, true = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
    X, W, y = (data[k].astype(np.float32) for k in ("site_covs", "obs_covs", "obs"))
    truth = np.concatenate([true["beta"][0], true["alpha"][0]])
    th = (truth + 2e-3 * np.random.default_rng(0).standard_normal((64, 10))).astype(np.float32)
    idx = list(range(0, 64, 8))
    ref_lp, ref_gr = c_oracle.occu_logp_grad(th[idx].astype(np.float64), X.astype(np.float64), W.astype(np.float64),
                                             y.astype(np.float64), dtype=np.float64)
    for name, env, kw in (("K1d", None, {}), ("K1c", "1", {}), ("strict", None, dict(strict_math=True))):
        if env: os.environ["BL_OCCU_CHAIN_KERNEL"] = env
        else: os.environ.pop("BL_OCCU_CHAIN_KERNEL", None)
        with bb.OccupancyLikelihood("occu", X, W, y, max_chains=64, **kw) as lk:
            lp, gr = lk.logp_and_grad(th)
        d = gr[idx].astype(np.float64) - ref_gr
        print(name, "logp err", np.round(lp[idx] - ref_lp, 3))
        print(name, "grad err mean over chains", np.round(d.mean(axis=0), 4), "std", np.round(d.std(axis=0), 4))
