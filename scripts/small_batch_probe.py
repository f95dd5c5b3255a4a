"""A few evaluations at C = 1 and C = 5 (config 2) for an ncu capture of the small-batch kernel (K1s, occu_small.cu;
BL_SMALL_KERNEL=0: the site-parallel engine)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=1_000_000, deployment_days_per_site=56)
for C in (1, 5):
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
        th = np.random.default_rng(C).uniform(-2, 2, size=(C, lk.theta_dim)).astype(np.float32)
        for _ in range(4):
            lk.logp_and_grad(th)
