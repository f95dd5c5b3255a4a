"""Error statistics of the default (bounded-error SFU) and strict (libm) fp32 kernels against the fp64
C oracle at config-2 size; printed as the table quoted in DESIGN.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
from oracle import c_oracle

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data, true = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
X, W, y = (data[k].astype(np.float32) for k in ("site_covs", "obs_covs", "obs"))
rng = np.random.default_rng(0)
truth = np.concatenate([true["beta"][0], true["alpha"][0]])
th = np.concatenate([rng.uniform(-2, 2, size=(120, 10)), truth + 2e-3 * rng.standard_normal((136, 10))]).astype(np.float32)
idx = list(range(0, 256, 8))
ref_lp, ref_gr = c_oracle.occu_logp_grad(th[idx].astype(np.float64), X.astype(np.float64), W.astype(np.float64),
                                         y.astype(np.float64), dtype=np.float64)
for name, kw in (("default (SFU ex2/lg2/rcp)", {}), ("BL_FLAG_STRICT_MATH (libm)", dict(strict_math=True))):
    with bb.OccupancyLikelihood("occu", X, W, y, max_chains=256, **kw) as lk:
        lp, gr = lk.logp_and_grad(th)
    lp, gr = lp[idx].astype(np.float64), gr[idx].astype(np.float64)
    rel = np.abs(lp - ref_lp) / np.abs(ref_lp)
    gscale = np.abs(ref_gr).max(axis=1, keepdims=True)
    gerr = np.abs(gr - ref_gr) / gscale
    near = np.array(idx) >= 120
    print(f"{name:28s} logp rel err max {rel.max():.2e} (abs max {np.abs(lp-ref_lp).max():.3f}); grad err/|g|inf: "
          f"random theta max {gerr[~near].max():.2e}, near mode max {gerr[near].max():.2e} "
          f"(abs max near mode {np.abs(gr-ref_gr)[near].max():.3f}, |g|inf near mode median {np.median(gscale[near]):.1f})")
