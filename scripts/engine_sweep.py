"""Small-chain-batch sweep of the site-parallel engine: us per evaluation vs C for the A/B switches
(BL_ENGINE_NCH = chains interleaved per pass, BL_ENGINE_BPS = resident blocks the ring is sized for,
BL_ENGINE_WC=old = previous warp arrangement).  Usage (GPU box): python scripts/engine_sweep.py [n_sites]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200.likelihood import DeviceBuffer

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56)
ref = {}
for wc, nch, bps in (("old", 1, 2), ("", 1, 2), ("", 2, 2), ("", 4, 2), ("", 4, 3), ("", 2, 3)):
    os.environ.pop("BL_ENGINE_WC", None)
    if wc:
        os.environ["BL_ENGINE_WC"] = wc
    os.environ["BL_ENGINE_NCH"] = str(nch)
    os.environ["BL_ENGINE_BPS"] = str(bps)
    row = []
    for C in (1, 2, 4, 5, 8, 16, 32, 64, 127):
        with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
            D = lk.theta_dim
            th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
            theta = np.random.default_rng(C).uniform(-2, 2, size=(C, D)).astype(np.float32)
            th.upload(theta)
            lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 20)
            ms = min(lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 200) for _ in range(3))
            out = lp.download((C,), np.float32), gr.download((C, D), np.float32)
            if C not in ref:
                ref[C] = out
            err = max(np.abs(out[0] - ref[C][0]).max() / np.abs(ref[C][0]).max(),
                      np.abs(out[1] - ref[C][1]).max() / np.abs(ref[C][1]).max())
            row.append(f"C={C}:{ms*1e3:.1f}us({err:.0e})")
            th.free(); lp.free(); gr.free()
    print(f"wc={wc or 'min':3s} nch={nch} bps={bps}  " + "  ".join(row), flush=True)
