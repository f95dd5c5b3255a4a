"""GPU: device-resident chain-batched NUTS vs the numpy NUTS restatement and vs the reference's
own recovery tests (biolith/models/occu.py:433-456, occu_rn.py:361-388, occu_cop.py:399-422)."""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_device_nuts_matches_cpu_nuts_posterior():
    import biolith_b200 as bb
    from biolith_b200 import diagnostics as dg
    from oracle import nuts as onuts
    from oracle import occupancy as orc

    g = load_golden("occu_default")
    d = g["data"]
    with bb.OccupancyLikelihood("occu", d["site_covs"], d["obs_covs"], d["obs"], max_chains=64) as lk:
        s = bb.NutsSampler(lk, 64, 400, 400, seed=7)
        assert s.run(timeout=300)
        res = s.results()
        s.close()
    x = res["samples"].astype(np.float64)  # (C, N, D)
    assert x.shape == (64, 400, 4) and np.all(res["n_saved"] == 400)
    assert np.all(dg.split_gelman_rubin(x) < 1.02)
    assert res["diverging"].mean() < 0.01
    assert 0.6 < res["accept_prob"].mean() < 0.95
    # CPU reference chains (same algorithm, numpy RNG)
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"])
    lpg = lambda th: orc.occu_logp_grad(th, pr)
    rng = np.random.default_rng(0)
    ref = np.stack([onuts.nuts_chain(lpg, rng.uniform(-2, 2, 4), 300, 400, np.random.default_rng(100 + i))["samples"]
                    for i in range(4)])
    m_gpu, m_ref = x.reshape(-1, 4).mean(0), ref.reshape(-1, 4).mean(0)
    se = np.sqrt(dg.mcse_mean(x) ** 2 + dg.mcse_mean(ref) ** 2)
    assert np.all(np.abs(m_gpu - m_ref) < 5 * se), (m_gpu, m_ref, se)
    sd_gpu, sd_ref = x.reshape(-1, 4).std(0), ref.reshape(-1, 4).std(0)
    assert np.all(np.abs(sd_gpu / sd_ref - 1) < 0.15)
    q_gpu = np.quantile(x.reshape(-1, 4), [0.05, 0.95], axis=0)
    q_ref = np.quantile(ref.reshape(-1, 4), [0.05, 0.95], axis=0)
    assert np.all(np.abs(q_gpu - q_ref) < 0.25 * sd_ref)


def test_fit_mirror_recovers_truth_like_reference_test():
    """The reference's own acceptance test (occu.py:433-456): psi-bar vs z-bar atol 0.1, coefs atol 0.5."""
    import biolith_b200 as bb

    data, true = bb.simulate_occupancy("occu", random_seed=0)
    res = bb.fit(bb.models.occu, **data, num_chains=8, num_samples=300, num_warmup=300, timeout=300)
    s = res.samples
    assert np.allclose(s["psi"].mean(), true["z"].mean(), atol=0.1)
    for i in range(2):
        assert np.allclose(s[f"cov_state_{i}"].mean(), true["beta"][0, i], atol=0.5)
        assert np.allclose(s[f"cov_det_{i}"].mean(), true["alpha"][0, i], atol=0.5)
    assert s["cov_state_0"].shape == (8 * 300, 1)
    summ = res.mcmc.summary()
    assert np.all(summ["beta"]["r_hat"] < 1.05)
    ss = res.mcmc.info["site_summary"]
    assert ss["psi_mean"].shape == (100, 1)
    assert np.allclose(ss["psi_mean"].mean(), s["psi"].mean(), atol=5e-3)
    # sites with a detection are occupied with certainty; the others less likely than a priori
    det = np.nansum(data["obs"][0, :, 0, :], axis=1) > 0
    assert np.all(ss["occupancy_prob"][det, 0] > 0.999) and np.all(ss["occupancy_prob"][~det, 0] < ss["psi_mean"][~det, 0])


def test_fit_rn_and_cop_recover_truth():
    import biolith_b200 as bb

    data, true = bb.simulate_occupancy("occu_rn", random_seed=0)
    res = bb.fit(bb.models.occu_rn, **data, num_chains=8, num_samples=250, num_warmup=250, timeout=600,
                 max_abundance=50)
    assert np.allclose(res.samples["abundance"].mean(), true["abundance"].mean(), rtol=0.25)
    data, true = bb.simulate_occupancy("occu_cop", random_seed=0, simulate_missing=True)
    res = bb.fit(bb.models.occu_cop, **data, num_chains=8, num_samples=250, num_warmup=250, timeout=600)
    # the cop posterior has a far-away local mode (psi -> 0, every count explained by the fp rate) that
    # traps an occasional U(-2,2) start -- the CPU NUTS restatement shows the same (1 chain in 6 at
    # lp -16925 vs -7153) -- so judge the bulk of the chains, as the reference's 5-chain mean does
    g = res.mcmc.get_samples(group_by_chain=True)
    pe = res.mcmc.get_extra_fields(group_by_chain=True)["potential_energy"].mean(axis=1)
    good = pe < np.median(pe) + 50.0
    assert good.sum() >= 6
    assert np.allclose(g["psi"][good].mean(), true["z"].mean(), atol=0.1)
    assert np.allclose(g["alpha"][good][..., 0].mean(), true["alpha"][0, 0], atol=0.5)
    assert np.allclose(g["beta"][good][..., 1].mean(), true["beta"][0, 1], atol=0.5)
    assert "rate_fp_constant" in res.samples


def test_unsupported_options_raise():
    import biolith_b200 as bb

    data, _ = bb.simulate_occupancy("occu", random_seed=0)
    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu, **data, coords=np.zeros((100, 2)))
    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu_rn, **data, site_random_effects=True)  # random effects: occu only
    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu, **data, kernel="hmc")


def test_fit_multi_season_and_multi_species_like_reference_tests():
    """biolith/models/occu.py:459-492: multi-period recovery (atol 0.15) and multi-species shapes."""
    import biolith_b200 as bb

    data, true = bb.simulate_occupancy("occu", simulate_missing=True, n_periods=3, random_seed=0)
    res = bb.fit(bb.models.occu, **data, num_chains=4, num_samples=300, num_warmup=300, timeout=600)
    assert res.samples["psi"].shape[1:] == (3, 100, 1)
    assert np.allclose(res.samples["psi"].mean(), true["z"].mean(), atol=0.15)
    data, _ = bb.simulate_occupancy("occu", simulate_missing=True, n_species=2, n_sites=30, random_seed=0)
    res = bb.fit(bb.models.occu, **data, num_chains=2, num_samples=100, num_warmup=100, timeout=600)
    assert res.samples["psi"].shape[-1] == 2
    assert res.samples["cov_state_0"].shape == (200, 2)
    # a shared false-positive parameter couples the species: one joint run over the composite handle
    res = bb.fit(bb.models.occu, **data, num_chains=2, num_samples=50, num_warmup=50, false_positives_constant=True)
    assert res.samples["cov_state_0"].shape == (100, 2) and res.samples["prob_fp_constant"].shape == (100,)


def test_fit_nmixture_recovers_truth():
    """The reference's own acceptance test, biolith/models/nmixture.py:377-420 (same simulator settings,
    max_abundance = largest observed count, abundance rtol 0.2, coefficients atol 0.5)."""
    import biolith_b200 as bb

    data, true = bb.simulate_occupancy("nmixture", simulate_missing=True, deployment_days_per_site=70,
                                       session_duration=7, min_abundance=1.0, min_observation_rate=1.0,
                                       max_observation_rate=6.0, random_seed=0)
    K = int(np.nanmax(data["obs"]))
    res = bb.fit(bb.models.nmixture, **data, max_abundance=K, num_chains=8, num_samples=300, num_warmup=300,
                 timeout=600)
    assert np.allclose(res.samples["abundance"].mean(), true["abundance"].mean(), rtol=0.2)
    for i in range(2):
        assert np.allclose(res.samples[f"cov_state_{i}"].mean(), true["beta"][0, i], atol=0.5)
        assert np.allclose(res.samples[f"cov_det_{i}"].mean(), true["alpha"][0, i], atol=0.5)
    ss = res.mcmc.info["site_summary"]
    N_true = true["N"][0, 0, :]
    assert np.corrcoef(ss["abundance_posterior_mean"][:, 0], N_true)[0, 1] > 0.8
    assert np.all(res.mcmc.summary()["beta"]["r_hat"] < 1.05)


def test_find_heuristic_step_size_option():
    """numpyro's HMC(find_heuristic_step_size=True) analogue: same posterior, different warm-up start."""
    import biolith_b200 as bb
    from biolith_b200 import diagnostics as dg

    g = load_golden("occu_default")
    d = g["data"]
    with bb.OccupancyLikelihood("occu", d["site_covs"], d["obs_covs"], d["obs"], max_chains=32) as lk:
        out = []
        for flag in (False, True):
            s = bb.NutsSampler(lk, 32, 300, 300, seed=3, find_heuristic_step_size=flag)
            assert s.run(timeout=120)
            out.append(s.results()["samples"].astype(np.float64))
            s.close()
    a, b = out
    se = np.sqrt(dg.mcse_mean(a) ** 2 + dg.mcse_mean(b) ** 2)
    assert np.all(np.abs(a.reshape(-1, 4).mean(0) - b.reshape(-1, 4).mean(0)) < 5 * se)
    assert np.all(dg.split_gelman_rubin(b) < 1.03)


def test_map_init_finds_the_mode_and_fit_accepts_it():
    import biolith_b200 as bb
    from biolith_b200.optim import find_map

    data, true = bb.simulate_occupancy("occu", n_site_covs=2, n_obs_covs=1, n_sites=20000,
                                       deployment_days_per_site=56, random_seed=2)
    with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"]) as lk:
        th, lp, info = find_map(lk)
        _, g = lk.logp_and_grad(th)
        # at the mode the gradient is tiny relative to its scale a posterior-sd away (~sqrt(n) = 140)
        assert np.abs(g).max() < 20.0
        assert np.allclose(th[:3], true["beta"][0], atol=0.1) and np.allclose(th[3:], true["alpha"][0], atol=0.1)
    res = bb.fit(bb.models.occu, **data, num_chains=16, num_samples=150, num_warmup=150, init_strategy="map")
    assert np.all(res.mcmc.summary()["beta"]["r_hat"] < 1.05)
    assert np.allclose(res.samples["cov_state_1"].mean(), true["beta"][0, 1], atol=0.1)


def test_fit_occu_cs_recovers_truth():
    """The reference's own acceptance test, biolith/models/occu_cs.py:365-392 (simulate_cs with missing
    data; psi atol 0.1, coefficients atol 0.5, score means / scales atol 1)."""
    import biolith_b200 as bb

    data, true = bb.simulate_occupancy("occu_cs", simulate_missing=True, random_seed=0)
    res = bb.fit(bb.models.occu_cs, **data, num_chains=8, num_samples=300, num_warmup=300, timeout=600)
    assert np.allclose(res.samples["psi"].mean(), true["z"].mean(), atol=0.1)
    for i in range(2):
        assert np.allclose(res.samples[f"cov_state_{i}"].mean(), true["beta"][0, i], atol=0.5)
        assert np.allclose(res.samples[f"cov_det_{i}"].mean(), true["alpha"][0, i], atol=0.5)
    for k in ("mu0", "mu1", "sigma0", "sigma1"):
        assert np.allclose(res.samples[k].mean(), true[k], atol=1), k
    assert np.all(res.samples["mu1"] > res.samples["mu0"])  # the truncation is a hard constraint
    s = res.mcmc.summary()
    assert np.all(s["beta"]["r_hat"] < 1.05) and np.all(s["sigma1"]["r_hat"] < 1.05)
    ss = res.mcmc.info["site_summary"]
    assert np.corrcoef(ss["occupancy_prob"][:, 0], true["z"][0, 0, :])[0, 1] > 0.9


@pytest.mark.parametrize("model,chains", [("occu", 5), ("occu_rn", 3)])
def test_staged_transition_kernel_is_bit_identical(monkeypatch, model, chains):
    """Few chains run the NUTS transition with one warp per chain on a shared-memory copy of the chain's state
    (nuts_advance_staged_kernel); it executes the same arithmetic in the same order as the thread-per-chain kernel,
    so with the same seed every draw, statistic and adapted step size must be identical."""
    import biolith_b200 as bb

    name = "occu_5x3" if model == "occu" else "rn_5x3"
    g = load_golden(name)
    d = g["data"]
    out = {}
    for staged in ("1", "0"):
        monkeypatch.setenv("BL_NUTS_STAGED", staged)
        kw = dict(max_abundance=g["model_kwargs"].get("max_abundance", 100)) if model == "occu_rn" else {}
        with bb.OccupancyLikelihood(model, d["site_covs"], d["obs_covs"], d["obs"], max_chains=chains, **kw) as lk:
            s = bb.NutsSampler(lk, chains, 150, 100, seed=11)
            assert s.run(timeout=300)
            out[staged] = s.results()
            s.close()
    a, b = out["1"], out["0"]
    assert np.all(a["n_saved"] == 100)
    for k in ("samples", "accept_prob", "num_steps", "diverging", "step_size", "leapfrogs"):
        assert np.array_equal(a[k], b[k]), f"{k} differs between the staged and the thread-per-chain transition kernel"
