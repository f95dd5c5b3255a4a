"""Batch-size sweep around the engine / lane=chain kernel crossover and the 128- vs 256-thread chain
blocks (BL_CHAIN_MIN, BL_CHAIN_VARIANT tuning switches).  Usage (GPU box): python scripts/chain_sweep.py [model]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biolith_b200 as bb
from biolith_b200.likelihood import DeviceBuffer

model = sys.argv[1] if len(sys.argv) > 1 else "occu"
S = {"occu": 1_000_000, "occu_rn": 200_000, "occu_cop": 500_000, "occu_cs": 500_000}[model]
days = {"occu": 56, "occu_rn": 70, "occu_cop": 84, "occu_cs": 70}[model]
data, _ = bb.simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=days)
kw = dict(max_abundance=50) if model == "occu_rn" else {}
fpc = bool(data.pop("false_positives_constant", False))
iters = 5 if model == "occu_rn" else 20
ref = {}
CONFIGS = (("engine", dict(BL_CHAIN_MIN="100000")), ("chain256", dict(BL_CHAIN_MIN="32", BL_CHAIN_VARIANT="3")),
           ("chain128", dict(BL_CHAIN_MIN="32", BL_CHAIN_VARIANT="2")), ("default", dict()))
if model == "occu_cs":  # block size follows the batch size only; compare engine and default
    CONFIGS = (("engine", dict(BL_CHAIN_MIN="100000")), ("default", dict()))
for label, env in CONFIGS:
    for k in ("BL_CHAIN_MIN", "BL_CHAIN_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    row = []
    for C in (32, 64, 96, 128, 192, 256, 384, 512, 1024):
        if label == "engine" and C > 128 and not (model == "occu_cs" and C in (256, 1024)):
            continue
        with bb.OccupancyLikelihood(model, data["site_covs"], data["obs_covs"], data["obs"], data.get("session_duration"),
                                    false_positives_constant=fpc, max_chains=C, **kw) as lk:
            D = lk.theta_dim
            th = DeviceBuffer(C * D * 4); lp = DeviceBuffer(C * 4); gr = DeviceBuffer(C * D * 4)
            theta = np.random.default_rng(C).uniform(-1, 1, size=(C, D)).astype(np.float32)
            if model == "occu_cs":  # score parameters near the generating values (where a sampler spends its time)
                theta[:, -4:] = (np.array([0.0, np.log(10.0), np.log(10.0), np.log(5.0)]) +
                                 0.1 * np.random.default_rng(1).standard_normal((C, 4))).astype(np.float32)
            th.upload(theta)
            lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, 2)
            ms = min(lk.eval_timed(th.ptr, C, lp.ptr, gr.ptr, 0, iters) for _ in range(2))
            out = lp.download((C,), np.float32), gr.download((C, D), np.float32)
            if C not in ref:
                ref[C] = out
            err = max(np.abs(out[0] - ref[C][0]).max() / np.abs(ref[C][0]).max(),
                      np.abs(out[1] - ref[C][1]).max() / np.abs(ref[C][1]).max())
            row.append(f"C={C}:{ms:.3f}ms({err:.0e})")
            th.free(); lp.free(); gr.free()
    print(f"{model} {label:9s} " + "  ".join(row), flush=True)
