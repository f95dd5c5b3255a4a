"""Property tests of the CPU oracle on random small shapes (hypothesis): the closed forms the kernels
implement must agree with the op-by-op enumerated restatement for every model, on shapes the fixtures do
not cover (multi-period, zero covariates, fully masked units, NaN covariates), and obey the invariances the
GPU tests rely on at full size (additivity over a site split, permutation of sites, masked-visit removal)."""

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import occupancy as orc

MODELS = ["occu", "occu_rn", "occu_cop", "nmixture", "occu_cs"]


def _problem(model, seed, S, P, J, ks, ko):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, P, J, ko))
    if model == "occu_cop":
        y = rng.poisson(2.0, size=(1, S, P, J)).astype(float)
    elif model == "nmixture":
        y = rng.binomial(rng.poisson(3.0, size=(1, S, P, 1)), 0.4, size=(1, S, P, J)).astype(float)
    elif model == "occu_cs":
        y = np.where(rng.uniform(size=(1, S, P, J)) < 0.3, rng.normal(10, 5, size=(1, S, P, J)),
                     rng.normal(0, 10, size=(1, S, P, J)))
    else:
        y = (rng.uniform(size=(1, S, P, J)) < 0.35).astype(float)
    y[rng.uniform(size=y.shape) < 0.25] = np.nan
    if S > 1:
        y[0, 0] = np.nan  # a fully masked site
    if ko:
        W[rng.uniform(size=W.shape) < 0.05] = np.nan
    if ks and S > 2:
        X[1, 0] = np.nan
    T = rng.uniform(0.5, 9.0, size=(S, P, J)) if model == "occu_cop" else None
    kw = {"occu_cop": dict(fp_constant=True), "occu_rn": dict(max_abundance=15),
          "nmixture": dict(max_abundance=25)}.get(model, {})
    D = ks + ko + 2 + orc.n_extras(model, **{k: v for k, v in kw.items() if k.startswith("fp_")})
    th = rng.uniform(-1.0, 1.0, size=D)
    if model == "occu_cs":
        th[-4:] = [0.2, np.log(9.0), np.log(9.0), np.log(5.5)]
    return X, W, y, T, kw, th


ENUM_KW = {"fp_constant": "false_positives_constant", "fp_unoccupied": "false_positives_unoccupied"}


@pytest.mark.parametrize("model", MODELS)
@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10_000), S=st.integers(1, 9), P=st.integers(1, 3), J=st.integers(1, 6),
       ks=st.integers(0, 3), ko=st.integers(0, 3), prior=st.booleans())
def test_closed_form_equals_enumerated_on_random_shapes(model, seed, S, P, J, ks, ko, prior):
    X, W, y, T, kw, th = _problem(model, seed, S, P, J, ks, ko)
    pr = orc.prepare(X, W, y, T, dtype=np.float64)
    lp, gr = orc.logp_grad(model, th, pr, dtype=np.float64, prior=prior, **kw)
    data = dict(site_covs=X, obs_covs=W, obs=y, session_duration=T)
    ref = orc.log_joint_enumerated(model, th, data, dtype=np.float64, prior=prior,
                                   **{ENUM_KW.get(k, k): v for k, v in kw.items()})
    assert np.isfinite(lp) and np.all(np.isfinite(gr))
    assert abs(lp - ref) <= 1e-9 * max(1.0, abs(ref)), (lp, ref)


@pytest.mark.parametrize("model", MODELS)
@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10_000), S=st.integers(2, 12), J=st.integers(1, 6), ks=st.integers(0, 3),
       ko=st.integers(0, 3))
def test_site_invariances(model, seed, S, J, ks, ko):
    X, W, y, T, kw, th = _problem(model, seed, S, 1, J, ks, ko)
    full = orc.logp_grad(model, th, orc.prepare(X, W, y, T, dtype=np.float64), dtype=np.float64, prior=False, **kw)

    def part(sl):
        Ts = None if T is None else T[sl]
        return orc.logp_grad(model, th, orc.prepare(X[sl], W[sl], y[:, sl], Ts, dtype=np.float64), dtype=np.float64,
                             prior=False, **kw)

    h = S // 2
    a, b = part(slice(0, h)), part(slice(h, S))
    np.testing.assert_allclose(a[0] + b[0], full[0], rtol=1e-11, atol=1e-11)      # additivity over a site split
    np.testing.assert_allclose(a[1] + b[1], full[1], rtol=1e-9, atol=1e-9)
    perm = np.random.default_rng(seed).permutation(S)
    p = part(perm)
    np.testing.assert_allclose(p[0], full[0], rtol=1e-11, atol=1e-11)             # permutation of sites
    np.testing.assert_allclose(p[1], full[1], rtol=1e-9, atol=1e-9)
    if model != "nmixture":  # (nmixture's truncation depends on the site's largest count: keep visits)
        # appending a visit that is masked must not change anything
        Wp = np.concatenate([W, np.zeros((S, 1, 1, ko))], axis=2)
        yp = np.concatenate([y, np.full((1, S, 1, 1), np.nan)], axis=3)
        Tp = None if T is None else np.concatenate([T, np.ones((S, 1, 1))], axis=2)
        q = orc.logp_grad(model, th, orc.prepare(X, Wp, yp, Tp, dtype=np.float64), dtype=np.float64, prior=False, **kw)
        np.testing.assert_allclose(q[0], full[0], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(q[1], full[1], rtol=1e-9, atol=1e-9)


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10_000), S=st.integers(1, 6), P=st.integers(1, 2), J=st.integers(1, 4),
       ks=st.integers(0, 2), ko=st.integers(0, 2), site_re=st.booleans(), obs_re=st.booleans(), prior=st.booleans())
def test_random_effects_groundwork(seed, S, P, J, ks, ko, site_re, obs_re, prior):
    """Checker for the next row (site / observation random effects of occu, occu.py:168-228): closed form =
    op-by-op enumerated form, elementwise gradients = finite differences, and it reduces to plain occu."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, P, J, ko))
    y = (rng.uniform(size=(1, S, P, J)) < 0.4).astype(float)
    y[rng.uniform(size=y.shape) < 0.2] = np.nan
    kw = dict(site_random_effects=site_re, obs_random_effects=obs_re, dtype=np.float64, prior=prior)
    D = orc.occu_re_dims(S, P, J, ks, ko, site_re, obs_re)
    th = 0.5 * rng.normal(size=D)
    pr = orc.prepare(X, W, y, dtype=np.float64)
    lp, g = orc.occu_re_logp_grad(th, pr, **kw)
    f = lambda t: orc.occu_re_log_joint_enumerated(t, X, W, y, **kw)  # noqa: E731
    assert abs(lp - f(th)) <= 1e-10 * max(1.0, abs(lp))
    fd = orc.finite_difference_grad(f, th, h=1e-6)
    assert np.abs(g - fd).max() <= 2e-6 * max(1.0, np.abs(g).max())
    if not (site_re or obs_re):
        ref_lp, ref_g = orc.occu_logp_grad(th, pr, dtype=np.float64, prior=prior)
        assert abs(lp - ref_lp) <= 1e-12 * max(1.0, abs(ref_lp))
        np.testing.assert_allclose(g, ref_g, rtol=1e-12, atol=1e-12)
