"""CPU: the fit() mirror rejects everything outside the accelerated path *before* touching a GPU
(no silent fallback), with the reference's keyword surface (biolith/utils/fit.py:16-32, occu.py:19-40)."""

import numpy as np
import pytest


def _data():
    rng = np.random.default_rng(0)
    return dict(site_covs=rng.normal(size=(6, 1)), obs_covs=rng.normal(size=(6, 1, 3, 1)),
                obs=(rng.uniform(size=(1, 6, 1, 3)) < 0.5).astype(float))


@pytest.mark.parametrize("kw", [
    dict(kernel="hmc"), dict(kernel="discrete_hmc_gibbs"),
    dict(site_random_effects=True, false_positives_constant=True),  # random effects: occu without fp extras only
    dict(coords=np.zeros((6, 2))), dict(init_strategy=lambda *a: None),
    dict(init_strategy="median"),
])
def test_options_outside_the_path_raise(kw):
    import biolith_b200 as bb

    with pytest.raises(bb.BiolithB200Error):
        bb.fit(bb.models.occu, **_data(), **kw)


def test_random_effects_keywords():
    from biolith_b200 import BiolithB200Error
    from biolith_b200.models import model_options

    class HalfNormal:
        def __init__(self, scale):
            self.scale = scale

    _, _, _, kw = model_options("occu", dict(site_random_effects=True, prior_site_re_sd=HalfNormal(0.5)))
    assert kw == {"site_random_effects": True, "prior_site_re_sd_scale": 0.5}
    _, _, _, kw = model_options("occu", dict(obs_random_effects=True, site_random_effects=True))
    assert kw == {"site_random_effects": True, "obs_random_effects": True}
    for model, k in (("occu_rn", dict(site_random_effects=True)), ("occu_cop", dict(obs_random_effects=True)),
                     ("occu", dict(site_random_effects=True, prior_site_re_sd=object()))):
        with pytest.raises(BiolithB200Error):
            model_options(model, k)


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


class _D:
    """Duck-typed numpyro distribution (class name + attributes is all fit() reads)."""

    def __init__(self, kind, **kw):
        self.__class__ = type(kind, (), {})
        self.__dict__.update(kw)


def test_unknown_and_inconsistent_keywords_raise():
    import biolith_b200 as bb

    for kw in (dict(false_positive_constant=True),             # typo -> unknown key, not a silent default
               dict(max_abundance=10),                          # not a keyword of occu (occu.py:19-40)
               dict(prior_rate_fp_constant=_D("Exponential", rate=1.0)),  # occu_cop's keyword, not occu's
               dict(false_positives_constant=True, false_positives_unoccupied=True)):  # occu.py:127-129
        with pytest.raises(bb.BiolithB200Error):
            bb.fit(bb.models.occu, **_data(), **kw)


def test_false_positive_priors_are_mapped_or_rejected():
    from biolith_b200 import BiolithB200Error
    from biolith_b200.models import model_options

    fpc, fpu, K, pk = model_options("occu", dict(false_positives_constant=True,
                                                 prior_prob_fp_constant=_D("Beta", concentration1=1.0, concentration0=3.0)))
    assert (fpc, fpu, pk) == (True, False, {"prior_fp_beta": (1.0, 3.0)})
    _, _, _, pk = model_options("occu_cop", dict(false_positives_unoccupied=True,
                                                 prior_rate_fp_unoccupied=_D("Exponential", rate=2.5)))
    assert pk == {"prior_fp_rate": 2.5}
    # a prior for a switched-off parameter is inert, exactly as in the reference
    assert model_options("occu", dict(prior_prob_fp_constant=_D("Uniform", low=0, high=1)))[3] == {}
    for model, kw in (("occu", dict(false_positives_constant=True, prior_prob_fp_constant=_D("Uniform", low=0, high=1))),
                      ("occu_cop", dict(false_positives_constant=True, prior_rate_fp_constant=_D("Gamma", concentration=2, rate=1))),
                      ("occu", dict(prior_beta=_D("Cauchy", loc=0.0, scale=1.0)))):
        with pytest.raises(BiolithB200Error):
            model_options(model, kw)
    assert model_options("occu_rn", dict(max_abundance=40, prior_alpha=_D("Normal", loc=0.5, scale=2.0)))[2:] == (
        40, {"prior_alpha": (0.5, 2.0)})


def _frames():
    import pandas as pd

    rng = np.random.default_rng(3)
    sites = [f"s{i}" for i in range(5)]
    site_covs = pd.DataFrame(rng.normal(size=(5, 2)), index=sites, columns=["elev", "forest"])
    cols = pd.MultiIndex.from_product([["temp", "wind"], [0, 1, 2]])
    obs_covs = pd.DataFrame(rng.normal(size=(5, 6)), index=sites, columns=cols)
    obs = pd.DataFrame((rng.uniform(size=(5, 3)) < 0.5).astype(float), index=sites)
    return site_covs, obs_covs, obs


def test_prepare_data_aligns_frames_and_names_like_the_reference():
    """biolith/utils/data.py:9-142: common row order = the first frame's index, MultiIndex columns reshaped to
    (S, P, J, Ko), names from levels[0]."""
    from biolith_b200.data import prepare_data

    site_covs, obs_covs, obs = _frames()
    perm = [3, 0, 4, 1, 2]
    X, W, y, T, sn, on = prepare_data(site_covs.iloc[perm], obs_covs.iloc[::-1], obs, None)
    assert sn == ["intercept", "elev", "forest"] and on == ["intercept", "temp", "wind"]
    np.testing.assert_array_equal(X, site_covs.to_numpy())            # re-ordered to obs's index
    assert W.shape == (5, 1, 3, 2) and y.shape == (5, 1, 3)
    np.testing.assert_array_equal(W[:, 0, :, 0], obs_covs["temp"].to_numpy())
    np.testing.assert_array_equal(W[:, 0, :, 1], obs_covs["wind"].to_numpy())
    # three column levels: (covariate, period, replicate)
    import pandas as pd

    cols3 = pd.MultiIndex.from_product([["temp", "wind"], [2020, 2021], [0, 1, 2]])
    oc3 = pd.DataFrame(np.arange(5 * 12, dtype=float).reshape(5, 12), index=obs.index, columns=cols3)
    W3 = prepare_data(site_covs, oc3, obs, None)[1]
    assert W3.shape == (5, 2, 3, 2)
    np.testing.assert_array_equal(W3[:, 1, :, 0], oc3["temp"][2021].to_numpy())
    with pytest.raises(ValueError):
        prepare_data(site_covs, obs_covs.droplevel(1, axis=1), obs, None)  # flat columns are rejected
    # plain arrays: period dimension inserted, default names
    out = prepare_data(np.zeros((4, 2)), np.zeros((4, 3)), np.zeros((4, 3)), np.ones((4, 3)))
    assert out[1].shape == (4, 1, 3, 1) and out[2].shape == (4, 1, 3) and out[3].shape == (4, 1, 3)
    assert out[4] == ["0", "1", "2"] and out[5] == ["0", "1"]


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/biolith"), reason="needs the reference checkout")
def test_prepare_data_against_the_reference_implementation(tmp_path):
    """Run the reference's own prepare_data (jnp -> numpy stand-in) on the same frames in a subprocess."""
    import pickle
    import subprocess
    import sys

    from biolith_b200.data import prepare_data

    site_covs, obs_covs, obs = _frames()
    args = (site_covs.iloc[[3, 0, 4, 1, 2]], obs_covs.iloc[::-1], obs, None)
    with open(tmp_path / "in.pkl", "wb") as fh:
        pickle.dump(args, fh)
    code = (
        "import sys, pickle, numpy as np; sys.path.insert(0, %r);"
        "from oracle import refshim; refshim.import_reference();"
        "from biolith.utils.data import prepare_data;"
        "a = pickle.load(open(%r, 'rb')); out = prepare_data(*a);"
        "pickle.dump([None if o is None else (np.asarray(o) if not isinstance(o, list) else o) for o in out], open(%r, 'wb'))"
    ) % (__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))),
         str(tmp_path / "in.pkl"), str(tmp_path / "out.pkl"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = pickle.load(open(tmp_path / "out.pkl", "rb"))
    mine = prepare_data(*args)
    for a, b in zip(mine, ref):
        if isinstance(b, list):
            assert [str(x) for x in a] == [str(x) for x in b]
        elif b is None:
            assert a is None
        else:
            np.testing.assert_array_equal(np.asarray(a), b)
