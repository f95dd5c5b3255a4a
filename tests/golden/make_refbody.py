"""Reference-BODY goldens: log-density and gradient obtained by EXECUTING the unmodified reference model
functions (biolith/models/{occu,occu_rn,occu_cop,nmixture,occu_cs}.py, regression/linear.py,
utils/modeling.py, utils/distributions.py under /root/reference) through oracle/refshim.py.

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_refbody.py

Writes tests/golden/<name>_refbody.npz next to every <name>.npz (same data, same thetas) plus a few
extra cases the simulators do not produce (two species sharing a false-positive parameter, two periods,
site / observation random effects) and the outputs of the reference's own closed forms
`log_likelihood_manual` / `lppd_manual` (biolith/evaluation/log_likelihood.py:55-98, lppd.py:64-106).
Nothing at test time or on the GPU box reads /root/reference: the tests only read the .npz files.

Keys:  ref_logp_{f32,f64}, ref_loglik_*, ref_grad_*, ref_gradlik_*   (f32 = data rounded to float32 and
numpyro's clamp_probs constants of float32, arithmetic in float64 -- the convention of the *_f32 oracle
goldens; f64 = everything float64), ref_logp_fp32arith (the body run op by op in numpy float32, value only).
"""

import ast
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refshim  # noqa: E402


def _data(g, dtype):
    d = {k: np.asarray(g[k], dtype).astype(np.float64) for k in ("site_covs", "obs_covs", "obs")}
    if "session_duration" in g:
        d["session_duration"] = np.asarray(g["session_duration"], np.float64)
    return d


def body_values(fn, model, thetas, g, kw):
    out = {}
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        data = _data(g, dt)
        with refshim.precision(np.float64, dt):
            rows = [refshim.value_and_grad(fn, model, th, data, **kw) for th in thetas]
        out[f"ref_logp_{tag}"] = np.array([r[0] for r in rows])
        out[f"ref_loglik_{tag}"] = np.array([r[1] for r in rows])
        out[f"ref_grad_{tag}"] = np.stack([r[2] for r in rows])
        out[f"ref_gradlik_{tag}"] = np.stack([r[3] for r in rows])
    with refshim.precision(np.float32):
        out["ref_logp_fp32arith"] = np.array(
            [refshim.value_and_grad(fn, model, th, _data(g, np.float32), h=0, **kw)[0] for th in thetas])
    return out


def existing_goldens(fns):
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        name = os.path.basename(path)[:-4]
        if name.endswith("_refbody") or name.startswith("extra_"):
            continue
        g = dict(np.load(path))
        model = str(g["model"])
        kw = ast.literal_eval(str(g["model_kwargs"]))
        if kw.get("fp_constant") and kw.get("fp_unoccupied"):
            # the reference asserts the two flags are exclusive (occu.py:127-129, occu_cop.py:113-115)
            print(f"{name}: skipped, the reference rejects both false-positive flags at once")
            continue
        out = body_values(fns[model], model, g["thetas"], g, kw)
        np.savez_compressed(os.path.join(HERE, name + "_refbody.npz"), **out)
        print(f"{name}: ref_logp_f64[0]={out['ref_logp_f64'][0]:.9f}  vs oracle golden {g['logp_f64'][0]:.9f}")


def _random_data(rng, S, P, J, Ks, Ko, Sp=1, counts=False, missing=0.15):
    X = rng.standard_normal((S, Ks))
    W = rng.standard_normal((S, P, J, Ko))
    y = rng.poisson(1.5, (Sp, S, P, J)).astype(float) if counts else (rng.random((Sp, S, P, J)) < 0.35).astype(float)
    y[rng.random(y.shape) < missing] = np.nan
    W[rng.random(W.shape) < 0.03] = np.nan
    X[rng.random(X.shape) < 0.02] = np.nan
    y[:, -1] = np.nan  # one fully missing site
    return X, W, y


def extras(fns):
    """Cases beyond the simulators: params are passed in numpyro's own site layout."""
    rng = np.random.default_rng(20261017)
    out = {}

    def run(tag, fn, params_list, args, kw, clamp=np.float64):
        vals, grads = [], []
        for params in params_list:
            with refshim.precision(np.float64, clamp):
                tr = refshim.trace(fn, params, **args, **kw)
                lj = float(np.real(tr.log_density()))
                ll = float(np.real(tr.log_likelihood_marginal()))
                g = {}
                for name, v in params.items():
                    v = np.asarray(v, np.float64)
                    gv = np.zeros(v.shape)
                    for idx in np.ndindex(*v.shape) if v.ndim else [()]:
                        pc = {k: np.asarray(x, np.complex128).copy() for k, x in params.items()}
                        pc[name][idx] += 1e-30j
                        gv[idx] = np.imag(refshim.trace(fn, pc, **args, **kw).log_density()) / 1e-30
                    g[name] = gv
            vals.append((lj, ll))
            grads.append(g)
        out[f"{tag}__logp"] = np.array([v[0] for v in vals])
        out[f"{tag}__loglik"] = np.array([v[1] for v in vals])
        for name in params_list[0]:
            out[f"{tag}__param__{name}"] = np.stack([np.asarray(p[name], np.float64) for p in params_list])
            out[f"{tag}__grad__{name}"] = np.stack([g[name] for g in grads])
        for k, v in args.items():
            out[f"{tag}__data__{k}"] = v
        out[f"{tag}__kwargs"] = np.array(repr(kw))

    # (1) two species sharing one false-positive probability (occu.py:146-150 sampled outside the species plate)
    X, W, y = _random_data(rng, 40, 1, 6, 2, 2, Sp=2)
    plist = [dict(beta=rng.uniform(-2, 2, (2, 3)), alpha=rng.uniform(-2, 2, (2, 3)),
                  prob_fp_constant=np.array(rng.uniform(-3, -1))) for _ in range(3)]
    run("occu_sp2_fpc", fns["occu"], plist, dict(site_covs=X, obs_covs=W, obs=y), dict(false_positives_constant=True))
    # (2) two periods, three species, no extras
    X, W, y = _random_data(rng, 30, 2, 5, 3, 1, Sp=3)
    plist = [dict(beta=rng.uniform(-2, 2, (3, 4)), alpha=rng.uniform(-2, 2, (3, 2))) for _ in range(3)]
    run("occu_p2_sp3", fns["occu"], plist, dict(site_covs=X, obs_covs=W, obs=y), {})
    # (3) occu_cop, two species sharing the unoccupied false-positive rate, exposure given
    X, W, y = _random_data(rng, 30, 1, 6, 2, 2, Sp=2, counts=True)
    T = rng.uniform(0.5, 3.0, (30, 1, 6))
    plist = [dict(beta=rng.uniform(-1, 1, (2, 3)), alpha=rng.uniform(-1, 1, (2, 3)),
                  rate_fp_unoccupied=np.array(rng.uniform(-2, 0))) for _ in range(3)]
    run("cop_sp2_fpu", fns["occu_cop"], plist, dict(site_covs=X, obs_covs=W, obs=y, session_duration=T),
        dict(false_positives_unoccupied=True))
    # (4) site + observation random effects (occu.py:170-173,191-196,215-218), both clamp conventions
    S, P, J = 25, 2, 4
    X, W, y = _random_data(rng, S, P, J, 2, 1)
    plist = [dict(beta=rng.uniform(-2, 2, (1, 3)), alpha=rng.uniform(-2, 2, (1, 2)),
                  site_re_sd=np.array(rng.uniform(-1, 0.5)), obs_re_sd=np.array(rng.uniform(-1, 0.5)),
                  site_re_occ=0.7 * rng.standard_normal((S, 1)), site_re_det=0.7 * rng.standard_normal((S, 1)),
                  obs_re=0.5 * rng.standard_normal((J, P, S, 1))) for _ in range(3)]
    run("occu_re_both", fns["occu"], plist, dict(site_covs=X, obs_covs=W, obs=y),
        dict(site_random_effects=True, obs_random_effects=True))
    plist_s = [{k: v for k, v in p.items() if k not in ("obs_re_sd", "obs_re")} for p in plist]
    run("occu_re_site", fns["occu"], plist_s, dict(site_covs=X, obs_covs=W, obs=y), dict(site_random_effects=True))
    run("occu_re_site_f32clamp", fns["occu"], plist_s,
        dict(site_covs=X.astype(np.float32).astype(np.float64), obs_covs=W.astype(np.float32).astype(np.float64),
             obs=y), dict(site_random_effects=True), clamp=np.float32)
    np.savez_compressed(os.path.join(HERE, "extra_refbody.npz"), **out)
    print("extra_refbody:", {k: v for k, v in out.items() if k.endswith("__logp")})


def deterministic_and_manual(fns):
    """Deterministic sites (psi, prob_detection: occu.py:207,221) per 'draw' and the reference's own closed
    forms log_likelihood_manual / lppd_manual evaluated on them."""
    import importlib

    # from /root/reference, under the stand-ins (the package re-exports same-named functions, hence sys.modules)
    importlib.import_module("biolith.evaluation")
    rll = sys.modules["biolith.evaluation.log_likelihood"]
    rlppd = sys.modules["biolith.evaluation.lppd"]

    out = {}
    for name in ("occu_missing", "occu_5x3"):
        g = dict(np.load(os.path.join(HERE, name + ".npz")))
        data = _data(g, np.float64)
        psi, pdet = [], []
        for th in g["thetas"]:
            Ks, Ko = data["site_covs"].shape[1], data["obs_covs"].shape[3]
            tr = refshim.trace(fns["occu"], refshim.theta_to_params("occu", th, Ks, Ko), **data)
            psi.append(np.asarray(tr.deterministic["psi"]))
            pdet.append(np.asarray(tr.deterministic["prob_detection"]))
        post = dict(psi=np.stack(psi), prob_detection=np.stack(pdet))
        llm = rll.log_likelihood_manual(post, data)
        out[f"{name}__psi"] = post["psi"]                       # (draws, S, Sp)
        out[f"{name}__prob_detection"] = post["prob_detection"]  # (draws, J, P, S, Sp)
        out[f"{name}__log_lik_manual"] = llm                     # (draws, Sp, S, P, J)
        out[f"{name}__lppd_manual"] = np.array(rlppd.lppd_manual(post, data))
    np.savez_compressed(os.path.join(HERE, "manual_refbody.npz"), **out)
    print("manual_refbody:", {k: (v.shape, float(v.ravel()[0])) for k, v in out.items()})


def main():
    m = refshim.import_reference()
    fns = {"occu": m.occu, "occu_rn": m.occu_rn, "occu_cop": m.occu_cop, "nmixture": m.nmixture,
           "occu_cs": m.occu_cs}
    which = sys.argv[1:] or ["goldens", "extras", "manual"]
    if "goldens" in which:
        existing_goldens(fns)
    if "extras" in which:
        extras(fns)
    if "manual" in which:
        deterministic_and_manual(fns)
    refshim.uninstall()


if __name__ == "__main__":
    main()
