"""One site-parallel-engine evaluation of config 2 at a small chain batch, for ncu.
usage: engine_probe.py [C] [model]  (BL_CHAIN_MIN=100000 forces the engine above 32 chains)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = sys.argv[2] if len(sys.argv) > 2 else "occu"
data, _ = bb.simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, n_sites=1_000_000 if model == "occu" else 500_000,
                                deployment_days_per_site=56 if model == "occu" else 70)
with bb.OccupancyLikelihood(model, data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
    th = np.random.default_rng(0).uniform(-2, 2, size=(C, lk.theta_dim))
    for _ in range(3):
        lk.logp_and_grad(th)
print("done")
