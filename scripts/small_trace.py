"""Block timeline of K1s: %globaltimer stamps per block, summarised (needs scripts/build_trace_lib.sh first)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200 import _lib
from biolith_b200.likelihood import DeviceBuffer

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libbiolith_b200_trace.so')
lib = _lib.load()
data, truth = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=1_000_000, deployment_days_per_site=56)
mode = np.concatenate([truth["beta"][0], truth["alpha"][0]]).astype(np.float32)
for C in (1, 5, 8):
    for where in ("U(-2,2)", "near truth"):
        with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], max_chains=C) as lk:
            D = lk.theta_dim
            th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
            theta = np.random.default_rng(C).uniform(-2, 2, size=(C, D)).astype(np.float32)
            if where != "U(-2,2)":
                theta = (mode[None] + 1e-3 * theta).astype(np.float32)
            th.upload(theta)
            ms = lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 50)
            buf = np.zeros((148, 8), dtype=np.uint64)
            rc = lib.bl_debug_small_trace(buf.ctypes.data_as(ctypes.c_void_p), 148)
            t = buf[:, :5].astype(np.int64)
            t0 = t[:, 0].min()
            r = (t - t0) / 1e3
            q = lambda v: f"min {v.min():6.2f} med {np.median(v):6.2f} max {v.max():6.2f}"
            print(f"C={C} {where}: {ms * 1e3:.1f} us/eval (rc {rc}); us after the first block's entry:")
            print(f"   entry       {q(r[:, 0])}")
            print(f"   first tile  {q(r[:, 1])}   (after own entry: {q(r[:, 1] - r[:, 0])})")
            print(f"   warp0 done  {q(r[:, 2])}")
            print(f"   published   {q(r[:, 3])}")
            print(f"   exit        {q(r[:, 4])}")
            th.free(); lp.free(); gr.free()
