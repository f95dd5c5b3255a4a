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


def test_occu_cs_prior_surface():
    """occu_cs.py:29-30: prior_mu / prior_sigma may be one distribution or a pair; the kernels carry a single
    zero-centred Normal and a single Gamma, anything else raises before a GPU is touched."""
    import biolith_b200 as bb

    class Normal:
        def __init__(self, loc=0.0, scale=1.0):
            self.loc, self.scale = loc, scale

    class Gamma:
        def __init__(self, concentration, rate):
            self.concentration, self.rate = concentration, rate

    rng = np.random.default_rng(1)
    d = dict(site_covs=rng.normal(size=(6, 1)), obs_covs=rng.normal(size=(6, 1, 3, 1)),
             obs=rng.normal(size=(1, 6, 1, 3)))
    for kw in (dict(prior_mu=(Normal(0, 10), Normal(0, 10))), dict(prior_mu=Normal(1.0, 10.0)),
               dict(prior_sigma=Normal(0, 1)), dict(prior_sigma=(Gamma(5, 1), Gamma(5, 1)))):
        with pytest.raises(bb.BiolithB200Error):
            bb.fit(bb.models.occu_cs, **d, **kw)
    d2 = dict(d, obs=rng.normal(size=(2, 6, 1, 3)))  # shared score parameters couple the species
    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu_cs, **d2)
