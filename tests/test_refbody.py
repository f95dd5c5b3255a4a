"""The oracle against REFERENCE-BODY goldens (tests/golden/*_refbody.npz).

Those fixtures were produced by executing the unmodified reference model functions
(/root/reference/biolith/models/*.py + regression/linear.py + utils/{modeling,distributions}.py) under
oracle/refshim.py (tests/golden/make_refbody.py; complex-step derivatives of the executed body).  These tests
run everywhere (no /root/reference needed): they pin oracle/occupancy.py -- the checker of every GPU parity
test -- to reference lines at 1e-12.
"""

import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, GOLDEN_NAMES, load_golden
from oracle import occupancy as orc

REFBODY = [n for n in GOLDEN_NAMES if os.path.exists(os.path.join(GOLDEN_DIR, n + "_refbody.npz"))]
DT = {"f32": np.float32, "f64": np.float64}


def _ref(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + "_refbody.npz")))


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1.0))


def _grel(a, b):
    b = np.asarray(b)
    return np.max(np.abs(np.asarray(a) - b) / np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1.0))


def test_every_supported_golden_has_a_reference_body_fixture():
    # cop_both_fp is the one exception: the reference asserts the two flags are exclusive
    assert set(GOLDEN_NAMES) - set(REFBODY) == {"cop_both_fp"}


@pytest.mark.parametrize("mode", ["f32", "f64"])
@pytest.mark.parametrize("name", REFBODY)
def test_closed_form_oracle_equals_reference_body(name, mode):
    g, r = load_golden(name), _ref(name)
    d = g["data"]
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"], d.get("session_duration"), dtype=DT[mode])
    lp, gr = orc.logp_grad(g["model"], g["thetas"], pr, dtype=DT[mode], **g["model_kwargs"])
    lpl, grl = orc.logp_grad(g["model"], g["thetas"], pr, dtype=DT[mode], prior=False, **g["model_kwargs"])
    # occu_rn: the reference's own `1 - (1 - r)**N` (occu_rn.py:213) cancels for states near the clamp, the
    # closed form carries log(1 - p) exactly; the gap is the reference's rounding, documented in DESIGN.md 2
    vt, gt = (1e-5, 1e-3) if (g["model"] == "occu_rn" and mode == "f64") else (
        (5e-8, 1e-6) if g["model"] == "occu_rn" else (1e-12, 1e-11))
    assert _rel(lp, r[f"ref_logp_{mode}"]) < vt
    assert _rel(lpl, r[f"ref_loglik_{mode}"]) < vt
    assert _grel(gr, r[f"ref_grad_{mode}"]) < gt
    assert _grel(grl, r[f"ref_gradlik_{mode}"]) < gt
    # the committed oracle goldens (what the GPU tests read) are the same numbers
    assert _rel(g[f"logp_{mode}"], r[f"ref_logp_{mode}"]) < vt
    assert _grel(g[f"grad_{mode}"], r[f"ref_grad_{mode}"]) < gt


@pytest.mark.parametrize("mode", ["f32", "f64"])
@pytest.mark.parametrize("name", REFBODY)
def test_enumerated_oracle_equals_reference_body(name, mode):
    """The op-by-op restatement uses the reference's own formulation, so it matches to rounding for every
    model, occu_rn included."""
    g, r = load_golden(name), _ref(name)
    d = {k: np.asarray(v, DT[mode]).astype(np.float64) if k != "session_duration" else v
         for k, v in g["data"].items()}
    kw = {{"fp_constant": "false_positives_constant", "fp_unoccupied": "false_positives_unoccupied"}.get(k, k): v
          for k, v in g["model_kwargs"].items()}
    for i in (0, 2, 6):
        v = orc.log_joint_enumerated(g["model"], g["thetas"][i], d, dtype=DT[mode], **kw)
        assert abs(v - r[f"ref_logp_{mode}"][i]) <= 1e-12 * abs(v), (name, i)


@pytest.mark.parametrize("name", REFBODY)
def test_reference_fp32_arithmetic_is_no_closer_than_the_tolerance(name):
    """The body run op by op in numpy float32 (the reference's default dtype) sits 1e-8 .. 1e-6 from the
    float64 truth: the 1e-5 GPU tolerance is measured against the truth, not against another fp32 rounding."""
    r = _ref(name)
    assert _rel(r["ref_logp_fp32arith"], r["ref_logp_f32"]) < 2e-6


def _extra():
    return dict(np.load(os.path.join(GOLDEN_DIR, "extra_refbody.npz")))


def _normal_lp(x):
    return float(np.sum(-0.5 * np.asarray(x) ** 2 - 0.5 * np.log(2 * np.pi)))


@pytest.mark.parametrize("tag,model,extra,fp", [("occu_sp2_fpc", "occu", "prob_fp_constant", dict(fp_constant=True)),
                                                ("occu_p2_sp3", "occu", None, {}),
                                                ("cop_sp2_fpu", "occu_cop", "rate_fp_unoccupied",
                                                 dict(fp_unoccupied=True))])
def test_species_are_independent_given_shared_extras(tag, model, extra, fp):
    """n_species > 1 (occu.py:182-186): one likelihood per species over the same covariates; a false-positive
    parameter is sampled once, outside the species plate (occu.py:146-157) and shared.  logp and every
    gradient of the executed reference body = sum of the single-species closed forms."""
    e = _extra()
    X, W, y = (e[f"{tag}__data__{k}"] for k in ("site_covs", "obs_covs", "obs"))
    T = e.get(f"{tag}__data__session_duration")
    Sp = y.shape[0]
    for i in range(e[f"{tag}__logp"].size):
        beta, alpha = e[f"{tag}__param__beta"][i], e[f"{tag}__param__alpha"][i]
        x = e[f"{tag}__param__{extra}"][i] if extra else None
        total, g_extra = 0.0, 0.0
        for sp in range(Sp):
            pr = orc.prepare(X, W, y[sp : sp + 1], T, dtype=np.float64)
            th = np.concatenate([beta[sp], alpha[sp]] + ([[x]] if extra else []))
            lp, gr = orc.logp_grad(model, th, pr, dtype=np.float64, prior=False, **fp)
            lpp, grp = orc.logp_grad(model, th, pr, dtype=np.float64, prior=True, **fp)
            total += lp + _normal_lp(beta[sp]) + _normal_lp(alpha[sp])
            kb = beta.shape[1]
            np.testing.assert_allclose(grp[:kb], e[f"{tag}__grad__beta"][i][sp], rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(grp[kb : kb + alpha.shape[1]], e[f"{tag}__grad__alpha"][i][sp], rtol=1e-10,
                                       atol=1e-10)
            if extra:
                g_extra += gr[-1]
                prior_extra, gprior_extra = lpp - lp - _normal_lp(beta[sp]) - _normal_lp(alpha[sp]), grp[-1] - gr[-1]
        if extra:
            total += prior_extra  # the shared parameter's prior + Jacobian counts once
            np.testing.assert_allclose(g_extra + gprior_extra, e[f"{tag}__grad__{extra}"][i], rtol=1e-10)
        np.testing.assert_allclose(total, e[f"{tag}__logp"][i], rtol=1e-12)


@pytest.mark.parametrize("tag,site,obs,dt", [("occu_re_both", True, True, np.float64),
                                             ("occu_re_site", True, False, np.float64),
                                             ("occu_re_site_f32clamp", True, False, np.float32)])
def test_random_effects_oracle_equals_reference_body(tag, site, obs, dt):
    """occu.py:170-173,191-196,215-218 executed vs oracle/occupancy.py:occu_re_logp_grad (elementwise grads)."""
    e = _extra()
    X, W, y = (e[f"{tag}__data__{k}"] for k in ("site_covs", "obs_covs", "obs"))
    S, P, J = X.shape[0], W.shape[1], W.shape[2]
    pr = orc.prepare(X, W, y, dtype=dt)
    for i in range(e[f"{tag}__logp"].size):
        p = {k.split("__param__")[1]: v[i] for k, v in e.items() if k.startswith(f"{tag}__param__")}
        g = {k.split("__grad__")[1]: v[i] for k, v in e.items() if k.startswith(f"{tag}__grad__")}
        parts = [p["beta"][0], p["alpha"][0]]
        gparts = [g["beta"][0], g["alpha"][0]]
        if site:
            parts.append([p["site_re_sd"]]); gparts.append([g["site_re_sd"]])
        if obs:
            parts.append([p["obs_re_sd"]]); gparts.append([g["obs_re_sd"]])
        if site:
            parts += [p["site_re_occ"][:, 0], p["site_re_det"][:, 0]]
            gparts += [g["site_re_occ"][:, 0], g["site_re_det"][:, 0]]
        if obs:  # numpyro layout (J, P, S, 1) -> oracle layout (S, P, J)
            parts.append(p["obs_re"][..., 0].transpose(2, 1, 0).ravel())
            gparts.append(g["obs_re"][..., 0].transpose(2, 1, 0).ravel())
        th = np.concatenate([np.atleast_1d(np.asarray(q, np.float64)) for q in parts])
        gref = np.concatenate([np.atleast_1d(np.asarray(q, np.float64)) for q in gparts])
        assert th.size == orc.occu_re_dims(S, P, J, X.shape[1], W.shape[3], site, obs)
        lp, gr = orc.occu_re_logp_grad(th, pr, site_random_effects=site, obs_random_effects=obs, dtype=dt)
        np.testing.assert_allclose(lp, e[f"{tag}__logp"][i], rtol=1e-12)
        np.testing.assert_allclose(gr, gref, rtol=1e-9, atol=1e-10)


def test_deterministic_sites_and_manual_closed_forms():
    """psi / prob_detection as the reference registers them (occu.py:207,221) and its own closed forms
    log_likelihood_manual (evaluation/log_likelihood.py:55-98) / lppd_manual (lppd.py:64-106), all executed
    from /root/reference, against the oracle's per-unit summaries."""
    m = dict(np.load(os.path.join(GOLDEN_DIR, "manual_refbody.npz")))
    for name in ("occu_missing", "occu_5x3"):
        g = load_golden(name)
        d = g["data"]
        pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"], dtype=np.float64)
        psi, pdet, llm = orc.occu_deterministic_sites(g["thetas"], pr, d["obs"])
        np.testing.assert_allclose(psi, m[f"{name}__psi"], rtol=1e-13)
        np.testing.assert_allclose(pdet, m[f"{name}__prob_detection"], rtol=1e-13)
        np.testing.assert_allclose(llm, m[f"{name}__log_lik_manual"], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(orc.lppd_manual(llm, d), m[f"{name}__lppd_manual"], rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/biolith"), reason="needs the reference checkout (build container)")
def test_fixture_regenerates_from_the_reference_checkout():
    """Re-execute the reference body for one fixture in a subprocess (the stand-ins live in sys.modules) and
    compare with the committed numbers: the .npz files are what make_refbody.py produces today."""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, %r);"
        "from oracle import refshim; m = refshim.import_reference();"
        "g = np.load(%r); r = np.load(%r);"
        "d = {k: g[k] for k in ('site_covs', 'obs_covs', 'obs')};"
        "v = [refshim.value_and_grad(m.occu, 'occu', th, d) for th in g['thetas'][:3]];"
        "assert np.allclose([x[0] for x in v], r['ref_logp_f64'][:3], rtol=1e-14);"
        "assert np.allclose(np.stack([x[2] for x in v]), r['ref_grad_f64'][:3], rtol=1e-12, atol=1e-12);"
        "print('ok')"
    ) % (os.path.dirname(GOLDEN_DIR.rstrip("/")).rsplit("/tests", 1)[0], os.path.join(GOLDEN_DIR, "occu_missing.npz"),
         os.path.join(GOLDEN_DIR, "occu_missing_refbody.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
