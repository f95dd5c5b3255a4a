"""Small-chain-batch sweep of the site-parallel engine: ms per evaluation vs C for BL_ENGINE_BPS settings.
Usage (GPU box): python scripts/engine_sweep.py [n_sites]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200.likelihood import DeviceBuffer

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
rng = np.random.default_rng(0)
for bps in (2, 3, 4, 6):
    os.environ["BL_ENGINE_BPS"] = str(bps)
    row = []
    for C in (1, 2, 4, 5, 8, 16, 32, 64, 127):
        with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
            D = lk.theta_dim
            th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
            th.upload(rng.uniform(-2, 2, size=(C, D)).astype(np.float32))
            lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 20)
            ms = min(lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 200) for _ in range(3))
            row.append(f"C={C}:{ms*1e3:.1f}us")
            th.free(); lp.free(); gr.free()
    print(f"bps={bps}  " + "  ".join(row), flush=True)
