"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference):  python tests/golden/make_golden.py

Inputs come from the *unmodified* reference simulators (imported with jax/numpyro stubbed,
see _refstub.py).  Outputs (logp, grad at fixed thetas) come from oracle/occupancy.py.  The
reference ships no golden vectors / KATs for this path, and numpyro/jax
are not installable here, so the *likelihood* values are an oracle regression pin, not a
reference-produced number ("parity unpinned"); the *data* side is
reference-produced and additionally pinned by the sha256 fingerprints below.
"""

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from _refstub import import_reference_models, remove_stubs  # noqa: E402

from oracle import occupancy as orc  # noqa: E402


def fingerprint(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()[:16]


# fingerprints recorded in SURVEY.md section 8c for the default-seed simulators
EXPECTED = {
    "occu_default": dict(site_covs="1c874d495328096e", obs_covs="0a7b24aca46f3070", obs="5a59a82c3a2d8d65"),
    "rn_default": dict(obs="41c72a953cac02eb"),
    "cop_default": dict(obs="98bb526f382112a4"),
}


def thetas_for(D, true_theta, seed):
    rng = np.random.default_rng(seed)
    th = [np.asarray(true_theta, np.float64), np.zeros(D)]
    th += list(rng.uniform(-2.0, 2.0, size=(4, D)))  # init_to_uniform(radius=2), fit.py:93
    th += [np.asarray(true_theta) + 0.05 * rng.standard_normal(D)]
    return np.stack(th)


def build(name, model, data, true, model_kw, sim_kw):
    fpc = bool(model_kw.get("fp_constant", False))
    fpu = bool(model_kw.get("fp_unoccupied", False))
    extras = []
    if model == "occu_cs":  # [mu0, log(mu1 - mu0), log sigma0, log sigma1] at the simulator's truth
        extras = [true["mu0"], np.log(true["mu1"] - true["mu0"]), np.log(true["sigma0"]), np.log(true["sigma1"])]
    elif model == "occu_cop":
        extras = ([np.log(0.1)] if fpc else []) + ([np.log(0.1)] if fpu else [])
    else:
        extras = ([-2.0] if fpc else []) + ([-2.0] if fpu else [])
    true_theta = np.concatenate([true["beta"][0], true["alpha"][0], extras])
    D = true_theta.size
    thetas = thetas_for(D, true_theta, seed=1234)
    out = dict(
        site_covs=data["site_covs"], obs_covs=data["obs_covs"], obs=data["obs"],
        thetas=thetas, model=np.array(model), sim_kwargs=np.array(repr(sim_kw)),
        model_kwargs=np.array(repr(model_kw)),
    )
    sd = data.get("session_duration")
    if sd is not None:
        out["session_duration"] = np.asarray(sd, dtype=np.float64)
    for mode, dt in (("f32", np.float32), ("f64", np.float64)):
        pr = orc.prepare(data["site_covs"], data["obs_covs"], data["obs"], sd, dtype=dt)
        lp, gr = orc.logp_grad(model, thetas, pr, dtype=dt, **model_kw)
        out[f"logp_{mode}"] = lp
        out[f"grad_{mode}"] = gr
        lpl, grl = orc.logp_grad(model, thetas, pr, dtype=dt, prior=False, **model_kw)
        out[f"loglik_{mode}"] = lpl
        out[f"gradlik_{mode}"] = grl
    out["mask"] = orc.expected_mask(data["site_covs"], data["obs_covs"], data["obs"])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    fps = {k: fingerprint(data[k]) for k in ("site_covs", "obs_covs", "obs")}
    for k, v in EXPECTED.get(name, {}).items():
        assert fps[k] == v, f"{name}.{k}: reference simulator fingerprint changed: {fps[k]} != {v}"
    print(f"{name}: D={D} S={data['obs'].shape[1]} J={data['obs'].shape[3]} "
          f"logp_f32[0]={out['logp_f32'][0]:.6f} size={os.path.getsize(path)/1024:.0f} KiB")


def main():
    m = import_reference_models()
    jobs = [
        # name, model, simulator, simulator kwargs, model kwargs
        ("occu_default", "occu", m.simulate, dict(), dict()),
        ("occu_missing", "occu", m.simulate, dict(simulate_missing=True, n_site_covs=2, n_obs_covs=2), dict()),
        ("occu_5x3", "occu", m.simulate,
         dict(n_site_covs=5, n_obs_covs=3, n_sites=300, deployment_days_per_site=56), dict()),
        ("occu_fp_const", "occu", m.simulate,
         dict(n_site_covs=2, n_obs_covs=1, n_sites=120, deployment_days_per_site=70, prob_fp_constant=0.05),
         dict(fp_constant=True)),
        ("occu_fp_unocc", "occu", m.simulate,
         dict(n_site_covs=2, n_obs_covs=1, n_sites=120, deployment_days_per_site=70, prob_fp_unoccupied=0.05,
              simulate_missing=True),
         dict(fp_unoccupied=True)),
        ("rn_default", "occu_rn", m.simulate_rn, dict(), dict(max_abundance=100)),
        ("rn_5x3", "occu_rn", m.simulate_rn,
         dict(n_site_covs=5, n_obs_covs=3, n_sites=150, deployment_days_per_site=70, simulate_missing=True),
         dict(max_abundance=50)),
        ("cop_default", "occu_cop", m.simulate_cop, dict(), dict(fp_constant=True)),
        ("cop_missing_5x3", "occu_cop", m.simulate_cop,
         dict(n_site_covs=5, n_obs_covs=3, n_sites=150, deployment_days_per_site=84, simulate_missing=True),
         dict(fp_constant=True)),
        ("nmix_default", "nmixture", m.simulate_nmixture, dict(), dict(max_abundance=100)),
        ("nmix_missing_5x3", "nmixture", m.simulate_nmixture,
         dict(n_site_covs=5, n_obs_covs=3, n_sites=150, deployment_days_per_site=70, simulate_missing=True),
         dict(max_abundance=60)),
        ("cs_default", "occu_cs", m.simulate_cs, dict(), dict()),
        ("cs_missing_5x3", "occu_cs", m.simulate_cs,
         dict(n_site_covs=5, n_obs_covs=3, n_sites=150, deployment_days_per_site=70, simulate_missing=True), dict()),
        ("cop_both_fp", "occu_cop", m.simulate_cop,
         dict(n_site_covs=1, n_obs_covs=2, n_sites=80, deployment_days_per_site=70),
         dict(fp_constant=True, fp_unoccupied=True)),
    ]
    only = sys.argv[1:]  # optional name prefixes: rebuild just those fixtures
    jobs = [j for j in jobs if not only or any(j[0].startswith(o) for o in only)]
    sims = [(name, model, sim(**sim_kw), sim_kw, model_kw) for name, model, sim, sim_kw, model_kw in jobs]
    remove_stubs()
    for name, model, (data, true), sim_kw, model_kw in sims:
        build(name, model, data, true, model_kw, sim_kw)


if __name__ == "__main__":
    main()
