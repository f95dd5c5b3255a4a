"""The repo's generator reproduces the reference simulators' arrays for the fixture seeds
(fixtures in tests/golden were produced by the *reference's* simulate*/ functions)."""

import numpy as np
import pytest

from biolith_b200.simulate import simulate_occupancy


@pytest.mark.parametrize("name", ["occu_default", "occu_missing", "occu_5x3", "occu_fp_const", "occu_fp_unocc",
                                  "rn_default", "rn_5x3", "cop_default", "cop_missing_5x3", "cop_both_fp",
                                  "nmix_default", "nmix_missing_5x3", "cs_default", "cs_missing_5x3"])
def test_generator_matches_reference_fixture(name):
    from conftest import load_golden

    g = load_golden(name)
    kw = dict(g["sim_kwargs"])
    if "prob_fp" in kw:
        kw["prob_fp_constant"] = kw.pop("prob_fp")
    data, _ = simulate_occupancy(g["model"], **kw)
    for k in ("site_covs", "obs_covs", "obs"):
        assert data[k].shape == g[k].shape
        assert np.array_equal(data[k], g[k], equal_nan=True), f"{name}.{k} differs from the reference simulator"
    if g["model"] == "occu_cop":
        assert np.array_equal(data["session_duration"], g["session_duration"])
