"""Small-batch engine: per-lane accumulation (BL_ENGINE_LANE=1, default) vs butterfly per (tile, chain) (=0).
us per evaluation at config 2 for the chain counts `fit` typically runs (reference default num_chains = 5)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200.likelihood import DeviceBuffer

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
ref = {}
for lane in ("1",):
    os.environ["BL_ENGINE_LANE"] = lane
    row = []
    for C in (1, 2, 4, 5, 8, 16, 31):
        with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
            D = lk.theta_dim
            th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
            theta = np.random.default_rng(C).uniform(-2, 2, size=(C, D)).astype(np.float32)
            th.upload(theta)
            lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 20)
            ms = min(lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 200) for _ in range(3))
            out = lp.download((C,), np.float32), gr.download((C, D), np.float32)
            ref.setdefault(C, out)
            err = max(np.abs(out[0] - ref[C][0]).max() / np.abs(ref[C][0]).max(),
                      np.abs(out[1] - ref[C][1]).max() / np.abs(ref[C][1]).max())
            row.append(f"C={C}:{ms*1e3:.1f}us({err:.0e})")
            th.free(); lp.free(); gr.free()
    print(f"lane={lane}  " + "  ".join(row), flush=True)
