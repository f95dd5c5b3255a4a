"""Probe: device NUTS on a config-2-shaped dataset; prints throughput / ESS statistics."""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb
from biolith_b200 import diagnostics as dg

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=1_000_000)
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--warmup", type=int, default=200)
ap.add_argument("--samples", type=int, default=100)
ap.add_argument("--timeout", type=float, default=300)
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--heuristic", action="store_true")
ap.add_argument("--map-init", action="store_true")
a = ap.parse_args()
data, true = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=a.sites, deployment_days_per_site=56)
X, W, y = (data[k].astype(np.float32) for k in ("site_covs", "obs_covs", "obs"))
lk = bb.OccupancyLikelihood("occu", X, W, y, max_chains=a.chains)
init = None
if a.map_init:
    from biolith_b200.optim import find_map, init_around
    t0 = time.perf_counter()
    th_map, lp_map, info = find_map(lk, verbose=True)
    print('map search', round(time.perf_counter() - t0, 2), 's  |grad|inf', info['grad_inf_norm'], 'theta', th_map.round(4).tolist())
    init = init_around(th_map, a.chains)
s = bb.NutsSampler(lk, a.chains, a.warmup, a.samples, seed=1, max_tree_depth=a.depth, init_params=init,
                   find_heuristic_step_size=a.heuristic)
t0 = time.perf_counter()
ok = s.run(timeout=a.timeout)
dt = time.perf_counter() - t0
r = s.results()
saved = r["n_saved"]
nmin = int(saved.min())
out = dict(complete=bool(ok), wall_s=dt, global_steps=int(r["global_steps"]), ms_per_step=1e3 * dt / max(r["global_steps"], 1),
           leapfrogs_mean=float(r["leapfrogs"].mean()), leapfrogs_max=int(r["leapfrogs"].max()),
           warmup_leapfrogs_mean=float(r["warmup_leapfrogs"].mean()), saved_min=nmin, saved_mean=float(saved.mean()),
           step_size_median=float(np.median(r["step_size"])), useful_frac=float(r["leapfrogs"].sum() / (a.chains * max(r["global_steps"], 1))))
if nmin >= 8:
    x = r["samples"][:, :nmin].astype(np.float64)
    ne = dg.effective_sample_size(x)
    out.update(ess_min=float(ne.min()), ess_median=float(np.median(ne)), rhat_max=float(dg.split_gelman_rubin(x).max()),
               accept_mean=float(r["accept_prob"][:, :nmin].mean()), steps_per_draw=float(r["num_steps"][:, :nmin].mean()),
               div_frac=float(r["diverging"][:, :nmin].mean()),
               post_mean=x.reshape(-1, x.shape[2]).mean(0).round(4).tolist(),
               truth=np.concatenate([true["beta"][0], true["alpha"][0]]).round(4).tolist())
    out["ess_min_per_s_total"] = out["ess_min"] / dt
print(json.dumps(out))
