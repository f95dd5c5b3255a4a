"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
usage: sanitize.py [model ...]   (default: all five)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biolith_b200 as bb

rng = np.random.default_rng(0)
ALL = {"occu": {}, "occu_rn": dict(max_abundance=12), "occu_cop": dict(false_positives_constant=True),
       "nmixture": dict(max_abundance=80), "occu_cs": {}}
todo = [m for m in sys.argv[1:] if m in ALL] or ([] if sys.argv[1:] else list(ALL))
S = int(os.environ.get("SANITIZE_SITES", "333"))
for model in todo:
    kw = ALL[model]
    data, _ = bb.simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, n_sites=S, deployment_days_per_site=56,
                                    simulate_missing=True, random_seed=1)
    T = data.get("session_duration")
    with bb.OccupancyLikelihood(model, data["site_covs"], data["obs_covs"], data["obs"], T, **kw) as lk:
        # site-parallel engine (1, 2 and 4 chains per pass), chain kernels with 128- and 256-thread blocks
        for C in (3, 7, 40, 130, 256):
            lp, gr = lk.logp_and_grad(rng.uniform(-1, 1, size=(C, lk.theta_dim)))
            assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
        lk.site_summary(rng.uniform(-1, 1, size=(5, lk.theta_dim)))
        s = bb.NutsSampler(lk, 130, 6, 4, seed=1)
        s.run(max_steps=200)
        s.close()
    print(model, "ok", flush=True)
if "small" in sys.argv[1:] or not sys.argv[1:]:
    # second half of round 2: K1s (register-decoded J = 8 and runtime-J quads, chain chunks, clamp fallback), the
    # cooperative / two-level split reductions of finish_block, the staged NUTS transition kernel
    for ks, ko, days in ((5, 3, 56), (2, 2, 35), (8, 4, 91)):
        data, _ = bb.simulate_occupancy("occu", n_site_covs=ks, n_obs_covs=ko, n_sites=S, deployment_days_per_site=days,
                                        simulate_missing=True, random_seed=3)
        with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"]) as lk:
            for C in (1, 5, 9, 31):
                assert lk.plan(C)["kernel"] == 7
                th = rng.uniform(-2, 2, size=(C, lk.theta_dim))
                th[0] *= 12.0  # visits at the clamps: the per-lane fallback
                lp, gr = lk.logp_and_grad(th)
                assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
            s = bb.NutsSampler(lk, 5, 6, 4, seed=1)  # warp-per-chain transition kernel
            s.run(max_steps=120)
            s.close()
    big = int(os.environ.get("SANITIZE_SITES_BIG", "40000"))  # >= 64 site splits: the two-level reduction
    data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=big, deployment_days_per_site=56,
                                    random_seed=4)
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"]) as lk:
        for C in (2, 256):
            pl = lk.plan(C)
            lp, gr = lk.logp_and_grad(rng.uniform(-1, 1, size=(C, lk.theta_dim)))
            assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
            print("  plan", C, pl, flush=True)
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], dtype="float64") as lk:
        pl = lk.plan(3)
        lp, gr = lk.logp_and_grad(rng.uniform(-1, 1, size=(3, lk.theta_dim)))
        assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
        print("  plan fp64", 3, pl, flush=True)
    print("small ok", flush=True)
if "extras" in sys.argv[1:] or not sys.argv[1:]:
    # round 2: random effects (K9), composite species (K10), per-observation log-likelihood (K11)
    data, _ = bb.simulate_occupancy("occu", n_site_covs=2, n_obs_covs=2, n_sites=S, n_species=2,
                                    deployment_days_per_site=42, simulate_missing=True, random_seed=2)
    one = data["obs"][:1]
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], one, site_random_effects=True,
                                obs_random_effects=True) as lk:
        lp, gr = lk.logp_and_grad(0.3 * rng.standard_normal((5, lk.theta_dim)))
        assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"],
                                false_positives_constant=True) as lk:
        for C in (3, 40):
            lp, gr = lk.logp_and_grad(rng.uniform(-1, 1, size=(C, lk.theta_dim)))
            assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], one) as lk:
        out = lk.pointwise_loglik(rng.uniform(-1, 1, size=(70, lk.theta_dim)))
        assert np.isfinite(out["lppd_total"])
    print("extras ok", flush=True)
