"""CPU: the fit() mirror rejects everything outside the accelerated path *before* touching a GPU
(no silent fallback), with the reference's keyword surface (biolith/utils/fit.py:16-32, occu.py:19-40)."""

import numpy as np
import pytest


def _data():
    rng = np.random.default_rng(0)
    return dict(site_covs=rng.normal(size=(6, 1)), obs_covs=rng.normal(size=(6, 1, 3, 1)),
                obs=(rng.uniform(size=(1, 6, 1, 3)) < 0.5).astype(float))


@pytest.mark.parametrize("kw", [
    dict(kernel="hmc"), dict(kernel="discrete_hmc_gibbs"), dict(site_random_effects=True),
    dict(obs_random_effects=True), dict(coords=np.zeros((6, 2))), dict(init_strategy=lambda *a: None),
    dict(init_strategy="median"),
])
def test_options_outside_the_path_raise(kw):
    import biolith_b200 as bb

    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu, **_data(), **kw)


def test_unknown_model_and_regressor_raise():
    import biolith_b200 as bb

    def occu_comb():
        pass

    with pytest.raises(bb.BiolithB200Error):
        bb.fit(occu_comb, **_data())

    class MLPRegression:
        pass

    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu, **_data(), regressor_det=MLPRegression)


def test_model_descriptors_mirror_reference_names():
    import biolith_b200 as bb

    assert set(bb.models.SUPPORTED) == {"occu", "occu_rn", "occu_cop", "nmixture", "occu_cs"}
    assert bb.models.occu.__name__ == "occu"
    with pytest.raises(RuntimeError):
        bb.models.occu()
