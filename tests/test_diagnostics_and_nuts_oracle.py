"""CPU tests: numpyro-definition diagnostics and the numpy NUTS restatement."""

import numpy as np

from biolith_b200 import diagnostics as dg
from oracle import nuts as onuts


def test_ess_of_iid_draws_is_about_n():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 2000, 3))
    ne = dg.effective_sample_size(x)
    assert ne.shape == (3,)
    assert np.all(ne > 0.8 * 8000) and np.all(ne < 1.25 * 8000)
    assert np.all(np.abs(dg.split_gelman_rubin(x) - 1.0) < 0.01)


def test_ess_of_ar1_matches_theory():
    rng = np.random.default_rng(1)
    rho, C, N = 0.9, 8, 20000
    x = np.zeros((C, N))
    e = rng.standard_normal((C, N))
    for t in range(1, N):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho**2) * e[:, t]
    ne = dg.effective_sample_size(x)
    theory = C * N * (1 - rho) / (1 + rho)
    assert abs(ne / theory - 1) < 0.15


def test_rhat_detects_disagreeing_chains():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 500))
    x[0] += 3.0
    assert dg.gelman_rubin(x) > 1.5
    s = dg.summary({"a": x})
    assert set(s["a"]) >= {"mean", "std", "median", "5.0%", "95.0%", "n_eff", "r_hat"}


def test_adaptation_schedule_matches_stan_windows():
    assert onuts.build_adaptation_schedule(1000) == [(0, 74), (75, 99), (100, 149), (150, 249), (250, 449),
                                                     (450, 949), (950, 999)]
    assert onuts.build_adaptation_schedule(10) == [(0, 9)]
    assert onuts.build_adaptation_schedule(100) == [(0, 14), (15, 89), (90, 99)]


def test_numpy_nuts_recovers_a_gaussian():
    rng = np.random.default_rng(3)
    mu = np.array([1.0, -2.0, 0.5])
    sd = np.array([0.5, 2.0, 0.1])

    def lpg(th):
        zz = (th - mu) / sd
        return -0.5 * np.dot(zz, zz), -zz / sd

    chains = [onuts.nuts_chain(lpg, rng.uniform(-2, 2, 3), 300, 400, np.random.default_rng(10 + i)) for i in range(4)]
    x = np.stack([c["samples"] for c in chains])
    ne = dg.effective_sample_size(x)
    mean, std = x.reshape(-1, 3).mean(0), x.reshape(-1, 3).std(0)
    assert np.all(np.abs(mean - mu) < 5 * sd / np.sqrt(ne))
    assert np.all(np.abs(std / sd - 1) < 0.15)
    assert np.all(dg.split_gelman_rubin(x) < 1.05)
    assert not any(c["diverging"].any() for c in chains)
