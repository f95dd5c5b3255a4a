"""occu with site / observation random effects (biolith/models/occu.py:170-173,191-196,215-218) on the GPU:
the log-density, the reduced gradients (beta, alpha, log sd) and the ELEMENTWISE gradients of every random
effect against the oracle (oracle/occupancy.py:occu_re_logp_grad) and against the executed reference body
(tests/golden/extra_refbody.npz), plus the reference's own acceptance assertions (occu.py:770-863)."""

import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _data(rng, S, P, J, Ks, Ko):
    X = rng.standard_normal((S, Ks))
    W = rng.standard_normal((S, P, J, Ko))
    y = (rng.random((1, S, P, J)) < 0.35).astype(float)
    y[rng.random(y.shape) < 0.15] = np.nan
    W[rng.random(W.shape) < 0.03] = np.nan
    if S > 4:
        X[rng.integers(0, S), rng.integers(0, Ks)] = np.nan
    return X, W, y


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("float64", 1e-10)])
@pytest.mark.parametrize("site,obs", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("S,P,J,Ks,Ko", [(37, 1, 5, 2, 1), (300, 2, 4, 3, 2), (1, 1, 3, 1, 1)])
def test_random_effects_against_oracle(S, P, J, Ks, Ko, site, obs, dtype, tol):
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(S + J)
    X, W, y = _data(rng, S, P, J, Ks, Ko)
    D = orc.occu_re_dims(S, P, J, Ks, Ko, site, obs)
    th = np.concatenate([rng.uniform(-1.5, 1.5, (5, Ks + Ko + 2)), rng.uniform(-1, 0.5, (5, int(site) + int(obs))),
                         0.6 * rng.standard_normal((5, D - (Ks + Ko + 2 + int(site) + int(obs))))], axis=1)
    npdt = np.float32 if dtype == "float32" else np.float64
    th = th.astype(npdt)
    pr = orc.prepare(X, W, y, dtype=npdt)
    for prior in (True, False):
        with bb.OccupancyLikelihood("occu", X, W, y, dtype=dtype, prior=prior, site_random_effects=site,
                                    obs_random_effects=obs) as lk:
            assert lk.theta_dim == D
            lp, gr = lk.logp_and_grad(th)
            lp2, gr2 = lk.logp_and_grad(th)
            assert np.array_equal(lp, lp2) and np.array_equal(gr, gr2), "deterministic"
        for i in range(len(th)):
            rl, rg = orc.occu_re_logp_grad(th[i].astype(np.float64), pr, site_random_effects=site,
                                           obs_random_effects=obs, dtype=npdt, prior=prior)
            assert abs(lp[i] - rl) <= tol * max(abs(rl), 1.0), (prior, i, lp[i], rl)
            # reduced entries relative to their scale, elementwise entries one by one (absolute + relative)
            ns = Ks + Ko + 2 + int(site) + int(obs)
            assert np.max(np.abs(gr[i, :ns] - rg[:ns])) <= tol * max(np.abs(rg[:ns]).max(), 1.0)
            np.testing.assert_allclose(gr[i, ns:], rg[ns:], rtol=50 * tol, atol=20 * tol)


@pytest.mark.parametrize("tag,site,obs,dtype", [("occu_re_both", True, True, "float64"),
                                                ("occu_re_site", True, False, "float64"),
                                                ("occu_re_site_f32clamp", True, False, "float32")])
def test_random_effects_against_the_executed_reference_body(tag, site, obs, dtype):
    import biolith_b200 as bb

    e = dict(np.load(os.path.join(GOLDEN_DIR, "extra_refbody.npz")))
    X, W, y = (e[f"{tag}__data__{k}"] for k in ("site_covs", "obs_covs", "obs"))
    tol = 1e-5 if dtype == "float32" else 1e-10
    with bb.OccupancyLikelihood("occu", X, W, y, dtype=dtype, prior=True, site_random_effects=site,
                                obs_random_effects=obs) as lk:
        for i in range(e[f"{tag}__logp"].size):
            p = {k.split("__param__")[1]: v[i] for k, v in e.items() if k.startswith(f"{tag}__param__")}
            g = {k.split("__grad__")[1]: v[i] for k, v in e.items() if k.startswith(f"{tag}__grad__")}
            order = ["beta", "alpha"] + (["site_re_sd"] if site else []) + (["obs_re_sd"] if obs else []) + (
                ["site_re_occ", "site_re_det"] if site else []) + (["obs_re"] if obs else [])

            def flat(d, k):
                v = np.asarray(d[k], np.float64)
                if k == "obs_re":  # numpyro (J, P, S, 1) -> site-major (S, P, J)
                    v = v[..., 0].transpose(2, 1, 0)
                return v.ravel()

            th = np.concatenate([flat(p, k) for k in order])
            gref = np.concatenate([flat(g, k) for k in order])
            lp, gr = lk.logp_and_grad(th)
            assert abs(lp - e[f"{tag}__logp"][i]) <= tol * abs(e[f"{tag}__logp"][i])
            np.testing.assert_allclose(gr, gref, rtol=100 * tol, atol=20 * tol * max(1.0, np.abs(gref).max()))


def test_fit_with_random_effects_meets_the_reference_assertions():
    """occu.py:770-863 (test_site_random_effects / test_combined_random_effects), shortened: sample sites exist with the
    reference's names and shapes, sd > 0, mean psi near the simulated occupancy (atol 0.15)."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy

    data, true = simulate_occupancy("occu", random_seed=0, simulate_missing=True, n_sites=50, deployment_days_per_site=84)
    res = bb.fit(bb.models.occu, data["site_covs"], data["obs_covs"], data["obs"], num_chains=4, num_warmup=200,
                 num_samples=100, site_random_effects=True, obs_random_effects=True, timeout=600)
    S, J = data["site_covs"].shape[0], data["obs_covs"].shape[2]
    smp = res.samples
    for k in ("site_re_sd", "site_re_occ", "site_re_det", "obs_re_sd", "obs_re"):
        assert k in smp, k
    assert smp["site_re_occ"].shape[1:] == (S, 1) and smp["obs_re"].shape[1:] == (J, 1, S, 1)
    assert smp["site_re_sd"].mean() > 0 and smp["obs_re_sd"].mean() > 0
    assert abs(smp["psi"].mean() - true["z"].mean()) < 0.15
    assert np.isfinite(smp["cov_state_0"]).all()
