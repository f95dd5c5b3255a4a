"""K1s (occu_small.cu) against the site-parallel engine for small chain batches: us per evaluation at config 2
(1M sites x 8 visits, Ks = 5, Ko = 3) and at a generic shape (Ks = 2, Ko = 2, J = 5), outputs compared with each other
and (first 3 chains, config 2) with the fp64 C oracle.  BL_SMALL_KERNEL is read when a handle plans a batch size."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200.likelihood import DeviceBuffer

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
CS = (1, 2, 4, 5, 8, 9, 16, 31)


def sweep(tag, data, oracle_chains=0, mode_theta=None, engine=True):
    outs = {}
    for small in (("0", "1") if engine else ("1",)):
        os.environ["BL_SMALL_KERNEL"] = small
        row = []
        for C in CS:
            with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
                D = lk.theta_dim
                th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
                theta = np.random.default_rng(C).uniform(-2, 2, size=(C, D)).astype(np.float32)
                if mode_theta is not None:
                    theta = (mode_theta[None, :] + 1e-3 * theta).astype(np.float32)  # within 2e-3 of the truth
                th.upload(theta)
                lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 20)
                ms = min(lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 200) for _ in range(3))
                outs[(small, C)] = (lp.download((C,), np.float32).copy(), gr.download((C, D), np.float32).copy(), theta)
                row.append(f"C={C}:{ms * 1e3:.1f}us")
                th.free(); lp.free(); gr.free()
        print(f"{tag} BL_SMALL_KERNEL={small}  " + "  ".join(row), flush=True)
    worst = 0.0
    for C in (CS if engine else ()):
        a, b = outs[("0", C)], outs[("1", C)]
        e = max(np.abs(a[0] - b[0]).max() / np.abs(a[0]).max(), np.abs(a[1] - b[1]).max() / np.abs(a[1]).max())
        worst = max(worst, e)
    print(f"{tag} K1s vs engine, worst relative difference over all C: {worst:.2e}", flush=True)
    if oracle_chains:
        from oracle import c_oracle
        lp1, gr1, theta = outs[("1", 5)]
        lp0, gr0, _ = outs.get(("0", 5), outs[("1", 5)])
        n = oracle_chains
        ref_lp, ref_gr = c_oracle.occu_logp_grad(theta[:n].astype(np.float64), data["site_covs"].astype(np.float64),
                                                 data["obs_covs"].astype(np.float64), data["obs"].astype(np.float64),
                                                 dtype=np.float64)
        for name, lp, gr in (("K1s", lp1, gr1), ("engine", lp0, gr0)):
            el = np.abs(lp[:n] - ref_lp).max() / np.abs(ref_lp).max()
            eg = np.abs(gr[:n] - ref_gr).max() / np.abs(ref_gr).max()
            print(f"{tag} {name} vs fp64 C oracle ({n} chains): logp {el:.2e}  grad/|g|inf {eg:.2e}", flush=True)


def f32(d):
    return {k: v.astype(np.float32) for k, v in d.items() if k in ("site_covs", "obs_covs", "obs")}


data, truth = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
data = f32(data)
sweep("config2 U(-2,2)", data, oracle_chains=3)
if len(sys.argv) > 2:  # quick: the first sweep only
    sys.exit(0)
sweep("config2 near truth", data, oracle_chains=3,
      mode_theta=np.concatenate([truth["beta"][0], truth["alpha"][0]]).astype(np.float32), engine=False)
data2, _ = bb.simulate_occupancy("occu", n_site_covs=2, n_obs_covs=2, n_sites=S, deployment_days_per_site=35,
                                 simulate_missing=True)
sweep("generic 2x2 J=5 missing", f32(data2))
