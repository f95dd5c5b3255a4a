"""The C/OpenMP restatement (CPU baseline) agrees with the numpy oracle on the golden fixtures."""

import numpy as np
import pytest

from conftest import load_golden


@pytest.mark.parametrize("name", ["occu_default", "occu_missing", "occu_5x3"])
def test_c_oracle_matches_numpy_oracle(name):
    from oracle import c_oracle

    g = load_golden(name)
    d = g["data"]
    lp, gr = c_oracle.occu_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"], dtype=np.float64)
    np.testing.assert_allclose(lp, g["logp_f64"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["grad_f64"], rtol=1e-9, atol=1e-9)
    lp, gr = c_oracle.occu_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"], dtype=np.float32)
    np.testing.assert_allclose(lp, g["logp_f32"], rtol=1e-5)
    scale = np.maximum(np.abs(g["grad_f32"]).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(gr - g["grad_f32"]) / scale).max() < 1e-5
    lp1, _ = c_oracle.occu_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"], dtype=np.float64,
                                     nthreads=1)
    np.testing.assert_allclose(lp1, lp * 0 + g["logp_f64"], rtol=1e-12)


@pytest.mark.parametrize("name", ["cop_default", "cop_missing_5x3", "cop_both_fp"])
def test_c_cop_oracle_matches_numpy_oracle(name):
    """The C restatement of occu_cop (BASELINE config 4; double arithmetic, clamp constants per dtype) is a
    third independent derivation next to the enumerated and the closed-form numpy ones."""
    from oracle import c_oracle

    g = load_golden(name)
    d, mk = g["data"], g["model_kwargs"]
    kw = dict(fp_constant=mk.get("fp_constant", False), fp_unoccupied=mk.get("fp_unoccupied", False))
    for mode, dt in (("f64", np.float64), ("f32", np.float32)):
        for prior, lk, gk in ((True, "logp", "grad"), (False, "loglik", "gradlik")):
            lp, gr = c_oracle.occu_cop_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"],
                                                 d.get("session_duration"), dtype=dt, prior=prior, **kw)
            ref_lp, ref_gr = g[f"{lk}_{mode}"], g[f"{gk}_{mode}"]
            np.testing.assert_allclose(lp, ref_lp, rtol=1e-12)
            scale = np.maximum(np.abs(ref_gr).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(gr - ref_gr) / scale).max() < 1e-11


@pytest.mark.parametrize("name", ["rn_default", "rn_5x3"])
def test_c_rn_oracle_matches_numpy_oracle(name):
    """C restatement of occu_rn (BASELINE config 3) against the numpy closed form on the goldens."""
    from oracle import c_oracle

    g = load_golden(name)
    d, mk = g["data"], g["model_kwargs"]
    for mode, dt in (("f64", np.float64), ("f32", np.float32)):
        for prior, lk, gk in ((True, "logp", "grad"), (False, "loglik", "gradlik")):
            lp, gr = c_oracle.occu_rn_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"],
                                                max_abundance=mk["max_abundance"], dtype=dt, prior=prior,
                                                fp_constant=mk.get("fp_constant", False))
            ref_lp, ref_gr = g[f"{lk}_{mode}"], g[f"{gk}_{mode}"]
            np.testing.assert_allclose(lp, ref_lp, rtol=1e-11)
            scale = np.maximum(np.abs(ref_gr).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(gr - ref_gr) / scale).max() < 1e-10


def test_c_rn_oracle_false_positive_constant():
    from oracle import c_oracle
    from oracle import occupancy as orc

    g = load_golden("rn_5x3")
    d = g["data"]
    rng = np.random.default_rng(0)
    th = np.concatenate([g["thetas"][:4], rng.uniform(-3, 0, size=(4, 1))], axis=1)
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"], dtype=np.float32)
    ref_lp, ref_gr = orc.logp_grad("occu_rn", th, pr, max_abundance=30, fp_constant=True, dtype=np.float32)
    lp, gr = c_oracle.occu_rn_logp_grad(th, d["site_covs"], d["obs_covs"], d["obs"], max_abundance=30,
                                        fp_constant=True)
    np.testing.assert_allclose(lp, ref_lp, rtol=1e-11)
    scale = np.maximum(np.abs(ref_gr).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(gr - ref_gr) / scale).max() < 1e-10


@pytest.mark.parametrize("name", ["nmix_default", "nmix_missing_5x3"])
def test_c_nmixture_oracle_matches_numpy_oracle(name):
    from oracle import c_oracle

    g = load_golden(name)
    d, mk = g["data"], g["model_kwargs"]
    for mode, dt in (("f64", np.float64), ("f32", np.float32)):
        for prior, lk, gk in ((True, "logp", "grad"), (False, "loglik", "gradlik")):
            lp, gr = c_oracle.nmixture_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"],
                                                 max_abundance=mk["max_abundance"], dtype=dt, prior=prior)
            ref_lp, ref_gr = g[f"{lk}_{mode}"], g[f"{gk}_{mode}"]
            np.testing.assert_allclose(lp, ref_lp, rtol=1e-11)
            scale = np.maximum(np.abs(ref_gr).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(gr - ref_gr) / scale).max() < 1e-10


@pytest.mark.parametrize("name", ["cs_default", "cs_missing_5x3"])
def test_c_occu_cs_oracle_matches_numpy_oracle(name):
    from oracle import c_oracle

    g = load_golden(name)
    d = g["data"]
    for mode, dt in (("f64", np.float64), ("f32", np.float32)):
        for prior, lk, gk in ((True, "logp", "grad"), (False, "loglik", "gradlik")):
            lp, gr = c_oracle.occu_cs_logp_grad(g["thetas"], d["site_covs"], d["obs_covs"], d["obs"], dtype=dt,
                                                prior=prior)
            ref_lp, ref_gr = g[f"{lk}_{mode}"], g[f"{gk}_{mode}"]
            np.testing.assert_allclose(lp, ref_lp, rtol=1e-11)
            scale = np.maximum(np.abs(ref_gr).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(gr - ref_gr) / scale).max() < 1e-10
