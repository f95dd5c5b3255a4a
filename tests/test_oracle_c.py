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
