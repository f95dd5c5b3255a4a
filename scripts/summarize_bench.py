import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        n = d.get("nuts") or {}
        print(d["config"]["workload"], d["config"]["sharding"], "N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3),
              "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              {k: (round(v, 3) if isinstance(v, float) else v) for k, v in n.items()
               if k in ("ess_per_sec_min", "accept_prob_mean", "step_size_median", "leapfrogs_per_draw", "ess_min", "wall_s", "r_hat_max", "ms_per_global_step", "useful_eval_frac", "chains")}, d.get("exchange_check"))
    else:
        print(ln.strip()[:200])
