"""Is the 8-shard site-sharded NUTS failure caused by every shard having its own true parameters?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
from biolith_b200 import diagnostics as dg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
parts = [bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=n, deployment_days_per_site=56, random_seed=r)
         for r in range(8)]
for name, sel in (("pooled 8 heterogeneous shards", range(8)), ("shard 0 only", [0])):
    X = np.concatenate([parts[r][0]["site_covs"] for r in sel]).astype(np.float32)
    W = np.concatenate([parts[r][0]["obs_covs"] for r in sel]).astype(np.float32)
    y = np.concatenate([parts[r][0]["obs"] for r in sel], axis=1).astype(np.float32)
    with bb.OccupancyLikelihood("occu", X, W, y, max_chains=256) as lk:
        s = bb.NutsSampler(lk, 256, 300, 100, seed=11)
        t0 = time.perf_counter(); ok = s.run(timeout=200); dt = time.perf_counter() - t0
        r = s.results(); s.close()
    x = r["samples"].astype(np.float64)
    print(name, "sites", X.shape[0], "ok", ok, "wall", round(dt, 1), "steps", r["global_steps"], "rhat_max",
          round(float(dg.split_gelman_rubin(x).max()), 3), "ess_min", round(float(dg.effective_sample_size(x).min())),
          "chain-mean spread", np.round(x.mean(axis=1).std(axis=0).max(), 4), "eps median", round(float(np.median(r["step_size"])), 4),
          "pe range", np.round(r["potential_energy"].mean(axis=1).min(), 1), np.round(r["potential_energy"].mean(axis=1).max(), 1))
