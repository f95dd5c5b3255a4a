"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb

rng = np.random.default_rng(0)
for model, kw in (("occu", {}), ("occu_rn", dict(max_abundance=20)), ("occu_cop", dict(false_positives_constant=True))):
    data, _ = bb.simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, n_sites=333, deployment_days_per_site=56,
                                    simulate_missing=True, random_seed=1)
    T = data.get("session_duration")
    with bb.OccupancyLikelihood(model, data["site_covs"], data["obs_covs"], data["obs"], T, **kw) as lk:
        for C in (3, 130):  # site-parallel engine, chain-parallel kernels
            lp, gr = lk.logp_and_grad(rng.uniform(-1, 1, size=(C, lk.theta_dim)))
            assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
        s = bb.NutsSampler(lk, 130, 6, 4, seed=1)
        s.run(max_steps=400)
        s.close()
    print(model, "ok")
