"""GPU parity tests proper: the CUDA path through the C ABI vs the oracle / committed goldens.

Tolerances (north star): fp32 -> 1e-5 relative on logp and on the gradient (relative to the
gradient's scale); fp64 -> 1e-10.  Masks must be bit-exact.
"""

import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

IMPLEMENTED = {"occu", "occu_rn", "occu_cop", "nmixture", "occu_cs"}
RTOL = {"float32": 1e-5, "float64": 1e-10}


def _make(g, dtype, prior=True, **kw):
    import biolith_b200 as bb

    d = g["data"]
    mk = g["model_kwargs"]
    return bb.OccupancyLikelihood(
        g["model"], d["site_covs"], d["obs_covs"], d["obs"], d.get("session_duration"),
        false_positives_constant=mk.get("fp_constant", False),
        false_positives_unoccupied=mk.get("fp_unoccupied", False),
        max_abundance=mk.get("max_abundance", 100), dtype=dtype, prior=prior, **kw)


def assert_close(lp, gr, lp_ref, gr_ref, rtol, what=""):
    lp, gr = np.asarray(lp, np.float64), np.asarray(gr, np.float64)
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr)), what
    rel = np.abs(lp - lp_ref) / np.maximum(np.abs(lp_ref), 1.0)
    assert rel.max() <= rtol, f"{what} logp rel err {rel.max():.3e} > {rtol}"
    scale = np.maximum(np.abs(gr_ref).max(axis=-1, keepdims=True), 1.0)
    gerr = (np.abs(gr - gr_ref) / scale).max()
    assert gerr <= rtol, f"{what} grad rel err {gerr:.3e} > {rtol}"


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_parity(name, dtype):
    g = load_golden(name)
    if g["model"] not in IMPLEMENTED:
        pytest.skip("model not built yet")
    mode = "f32" if dtype == "float32" else "f64"
    with _make(g, dtype) as lk:
        assert lk.theta_dim == g["thetas"].shape[1]
        lp, gr = lk.logp_and_grad(g["thetas"])
        assert_close(lp, gr, g[f"logp_{mode}"], g[f"grad_{mode}"], RTOL[dtype], f"{name}/{dtype}")
        assert np.array_equal(lk.mask(), g["mask"][0]), "NaN mask is not bit-exact"
    with _make(g, dtype, prior=False) as lk:
        lp, gr = lk.logp_and_grad(g["thetas"])
        assert_close(lp, gr, g[f"loglik_{mode}"], g[f"gradlik_{mode}"], RTOL[dtype], f"{name}/{dtype}/lik")


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if n != "cop_both_fp"])
def test_parity_against_the_executed_reference_body(name, dtype):
    """CUDA vs numbers produced by RUNNING the unmodified reference model functions (tests/golden/*_refbody.npz,
    made by tests/golden/make_refbody.py through oracle/refshim.py) -- no restated model body in between."""
    import os

    from conftest import GOLDEN_DIR

    g = load_golden(name)
    r = np.load(os.path.join(GOLDEN_DIR, name + "_refbody.npz"))
    mode = "f32" if dtype == "float32" else "f64"
    # occu_rn in fp64: the reference's own `1 - (1 - r)**N` (occu_rn.py:213) cancels near the fp64 clamp: its value
    # is off by 3.5e-6 and its gradient by 3.4e-4 at the U(-2,2) thetas of rn_5x3 (tests/test_refbody.py, DESIGN.md
    # 2); the kernels carry log(1 - p) exactly and stay within 1e-10 of the closed-form oracle (test_golden_parity)
    rtol = 1e-3 if (g["model"] == "occu_rn" and dtype == "float64") else RTOL[dtype]
    with _make(g, dtype) as lk:
        lp, gr = lk.logp_and_grad(g["thetas"])
        assert_close(lp, gr, r[f"ref_logp_{mode}"], r[f"ref_grad_{mode}"], rtol, f"{name}/{dtype}/refbody")
    with _make(g, dtype, prior=False) as lk:
        lp, gr = lk.logp_and_grad(g["thetas"])
        assert_close(lp, gr, r[f"ref_loglik_{mode}"], r[f"ref_gradlik_{mode}"], rtol, f"{name}/{dtype}/refbody/lik")


@pytest.mark.parametrize("n_chains", [1, 2, 3, 5, 8, 16, 31, 32, 33, 64, 96, 127, 128, 129, 192, 256, 257, 384, 700])
def test_chain_batching_is_consistent(n_chains):
    """Every (C -> WS x WC arrangement, chunking) must give the same per-chain numbers."""
    from oracle import occupancy as orc

    g = load_golden("occu_5x3")
    d = g["data"]
    rng = np.random.default_rng(n_chains)
    th = rng.uniform(-2, 2, size=(n_chains, 10))
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"])
    ref_lp, ref_gr = orc.logp_grad("occu", th[: min(n_chains, 12)], pr)
    with _make(g, "float32") as lk:
        lp, gr = lk.logp_and_grad(th)
        k = min(n_chains, 12)
        assert_close(lp[:k], gr[:k], ref_lp, ref_gr, 1e-5, f"C={n_chains}")
        # chain c of a batch == the same theta evaluated alone (bitwise: same reduction order)
        lp1, gr1 = lk.logp_and_grad(th[-1])
        np.testing.assert_allclose(lp1, lp[-1], rtol=2e-6)
        np.testing.assert_allclose(gr1, gr[-1], rtol=2e-5, atol=2e-4)
        # determinism: same call twice -> identical bits
        lp2, gr2 = lk.logp_and_grad(th)
        assert np.array_equal(lp, lp2) and np.array_equal(gr, gr2)


@pytest.mark.parametrize("S,P,J,ks,ko", [(1, 1, 1, 1, 1), (33, 1, 5, 2, 1), (257, 1, 8, 5, 3), (40, 3, 4, 3, 2),
                                         (1000, 1, 40, 1, 1), (64, 2, 33, 7, 9), (50, 1, 6, 0, 0)])
@pytest.mark.parametrize("model", ["occu", "occu_rn", "occu_cop", "nmixture", "occu_cs"])
def test_ragged_shapes_against_oracle(model, S, P, J, ks, ko):
    import biolith_b200 as bb
    from oracle import occupancy as orc

    if model not in IMPLEMENTED:
        pytest.skip("model not built yet")
    rng = np.random.default_rng(S * 7 + J)
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
    # ragged visits: trailing replicates missing, plus NaNs in covariates
    lens = rng.integers(0, J + 1, size=(S, P))
    y[0][np.arange(J)[None, None, :] >= lens[:, :, None]] = np.nan
    if ko:
        W[rng.uniform(size=W.shape) < 0.03] = np.nan
    if ks and S > 3:
        X[rng.integers(0, S), rng.integers(0, ks)] = np.nan
    T = rng.uniform(0.5, 9.0, size=(S, P, J)) if model == "occu_cop" else None
    kw = dict(fp_constant=True) if model == "occu_cop" else (
        dict(max_abundance=30) if model in ("occu_rn", "nmixture") else {})
    D = ks + ko + 2 + (1 if model == "occu_cop" else 4 if model == "occu_cs" else 0)
    th = rng.uniform(-1.5, 1.5, size=(6, D))
    if model == "occu_cs":
        th[0, -4:] = [0.3, np.log(9.0), np.log(11.0), np.log(4.0)]  # near the generating score distributions
    pr = orc.prepare(X, W, y, T)
    ref_lp, ref_gr = orc.logp_grad(model, th, pr, **kw)
    with bb.OccupancyLikelihood(model, X, W, y, T, false_positives_constant=(model == "occu_cop"),
                                max_abundance=30) as lk:
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp, gr, ref_lp, ref_gr, 1e-5, f"{model} S={S} P={P} J={J}")
        assert np.array_equal(lk.mask(), orc.expected_mask(X, W, y)[0])


def test_extreme_thetas_hit_the_clamps():
    """|nu| beyond the clamp thresholds: values saturate at log(eps)/log(tiny), gradients vanish."""
    from oracle import occupancy as orc

    g = load_golden("occu_default")
    d = g["data"]
    th = np.array([[30.0, 5.0, 25.0, -9.0], [-40.0, 3.0, -30.0, 4.0], [100.0, 0.0, 100.0, 0.0],
                   [-100.0, 0.0, -100.0, 0.0]])
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"])
    ref_lp, ref_gr = orc.logp_grad("occu", th, pr)
    with _make(g, "float32") as lk:
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp, gr, ref_lp, ref_gr, 1e-5, "extreme")


def test_rejects_non_binary_detections():
    import biolith_b200 as bb

    X = np.zeros((4, 1))
    W = np.zeros((4, 1, 3, 1))
    y = np.full((1, 4, 1, 3), 2.0)
    with pytest.raises(bb.BiolithB200Error):
        bb.OccupancyLikelihood("occu", X, W, y)


def test_large_shape_properties():
    """Config-2 shape (scaled to 200k sites to stay in seconds): size-independent properties.
    (1) additivity over a site split: logp(A u B) - prior = loglik(A) + loglik(B);
    (2) permutation invariance of sites; (3) masked observations do not change the result."""
    import biolith_b200 as bb

    rng = np.random.default_rng(5)
    S, J, ks, ko = 200_000, 8, 5, 3
    X = rng.normal(size=(S, ks)).astype(np.float32)
    W = rng.normal(size=(S, 1, J, ko)).astype(np.float32)
    y = (rng.uniform(size=(1, S, 1, J)) < 0.3).astype(np.float32)
    th = rng.uniform(-1, 1, size=(16, 10))
    kw = dict(prior=False)
    with bb.OccupancyLikelihood("occu", X, W, y, **kw) as full:
        lp, gr = full.logp_and_grad(th)
    h = S // 3
    with bb.OccupancyLikelihood("occu", X[:h], W[:h], y[:, :h], **kw) as a, \
            bb.OccupancyLikelihood("occu", X[h:], W[h:], y[:, h:], **kw) as b:
        la, ga = a.logp_and_grad(th)
        lb, gb = b.logp_and_grad(th)
    np.testing.assert_allclose(la.astype(np.float64) + lb, lp, rtol=2e-6)
    np.testing.assert_allclose(ga.astype(np.float64) + gb, gr, rtol=1e-5, atol=1e-5 * np.abs(gr).max())
    perm = rng.permutation(S)
    with bb.OccupancyLikelihood("occu", X[perm], W[perm], y[:, perm], **kw) as pm:
        lp2, gr2 = pm.logp_and_grad(th)
    np.testing.assert_allclose(lp2, lp, rtol=2e-6)
    np.testing.assert_allclose(gr2, gr, rtol=1e-5, atol=1e-5 * np.abs(gr).max())
    # append fully-masked sites: nothing may change except exact zeros added
    Xp = np.concatenate([X, np.full((1000, ks), np.nan, np.float32)])
    Wp = np.concatenate([W, rng.normal(size=(1000, 1, J, ko)).astype(np.float32)])
    yp = np.concatenate([y, np.ones((1, 1000, 1, J), np.float32)], axis=1)
    with bb.OccupancyLikelihood("occu", Xp, Wp, yp, **kw) as pad:
        lp3, gr3 = pad.logp_and_grad(th)
        assert pad.n_masked == 1000 * J
    # a fully masked site contributes logaddexp(log psi, log(1-psi)) = 0 up to rounding
    np.testing.assert_allclose(lp3, lp, rtol=2e-6)


def test_config2_full_size_against_c_oracle():
    """BASELINE.json configs[1] at full size: 1M sites x 8 visits, 5+3 covariates, 1024 chains.
    A sample of chains is checked against the fp64 C/OpenMP oracle; plus the checksum property
    sum_c logp_c is reproduced when the chains are evaluated in a different batch split."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy
    from oracle import c_oracle

    data, _ = simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=1_000_000,
                                 deployment_days_per_site=56, random_seed=0)
    X = data["site_covs"].astype(np.float32)
    W = data["obs_covs"].astype(np.float32)
    y = data["obs"].astype(np.float32)
    th = np.random.default_rng(3).uniform(-2, 2, size=(1024, 10)).astype(np.float32)
    with bb.OccupancyLikelihood("occu", X, W, y, max_chains=1024) as lk:
        lp, gr = lk.logp_and_grad(th)
        idx = [0, 1, 255, 256, 700, 1023]
        ref_lp, ref_gr = c_oracle.occu_logp_grad(th[idx].astype(np.float64), X.astype(np.float64),
                                                 W.astype(np.float64), y.astype(np.float64), dtype=np.float64)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, "config2 full size")
        lp_a, gr_a = lk.logp_and_grad(th[:300])   # different chunking (2 x 150 chains)
        lp_b, gr_b = lk.logp_and_grad(th[300:])
        np.testing.assert_allclose(np.concatenate([lp_a, lp_b]), lp, rtol=1e-6)
        np.testing.assert_allclose(np.concatenate([gr_a, gr_b]), gr, rtol=1e-5, atol=1e-5 * np.abs(gr).max())


@pytest.mark.parametrize("missing", [False, True])
def test_config2_full_size_near_the_mode(missing):
    """Config 2 (and its simulate_missing=True variant) with the chains where NUTS spends its time:
    within 2e-3 of the simulating truth, where |g|inf ~ 2e3 is tiny against sum |terms| ~ 3e5 and against
    |H| |theta| ~ 1e6 -- any fixed relative error in an internal constant shows up here first.  Both math modes
    must stay within 1e-5 of the fp64 C oracle (gradient relative to |g|inf)."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy
    from oracle import c_oracle

    data, true = simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=1_000_000,
                                    deployment_days_per_site=56, random_seed=0, simulate_missing=missing)
    X, W, y = (data[k].astype(np.float32) for k in ("site_covs", "obs_covs", "obs"))
    truth = np.concatenate([true["beta"][0], true["alpha"][0]])
    rng = np.random.default_rng(0)
    th = (truth + 2e-3 * rng.standard_normal((256, 10))).astype(np.float32)
    idx = list(range(0, 256, 16))
    ref_lp, ref_gr = c_oracle.occu_logp_grad(th[idx].astype(np.float64), X.astype(np.float64), W.astype(np.float64),
                                             y.astype(np.float64), dtype=np.float64, nthreads=0)
    assert np.abs(ref_gr).max() < 2e-2 * 3e5, "the chains are meant to sit near the mode"
    for kw in ({}, dict(strict_math=True)):
        with bb.OccupancyLikelihood("occu", X, W, y, max_chains=256, **kw) as lk:
            from oracle import occupancy as orc

            assert np.array_equal(lk.mask(), orc.expected_mask(X, W, y)[0]), "NaN mask is not bit-exact at full size"
            lp, gr = lk.logp_and_grad(th)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"config2 near mode missing={missing} {kw}")


@pytest.mark.parametrize("model,sim_kw,model_kw,okw", [
    ("occu_rn", dict(n_sites=200_000, deployment_days_per_site=70), dict(max_abundance=50), dict(max_abundance=50)),
    ("occu_cop", dict(n_sites=500_000, deployment_days_per_site=84, simulate_missing=True),
     dict(false_positives_constant=True), dict(fp_constant=True)),
])
def test_configs_3_and_4_full_size_against_c_oracle(model, sim_kw, model_kw, okw):
    """BASELINE configs[2] (occu_rn 200k x 10, K = 50) and configs[3] (occu_cop 500k x 12, NaN-masked) at FULL size:
    a sample of the chains of a 256-chain batch (the lane = chain kernels) against the C/OpenMP oracle in double
    arithmetic, at U(-2,2) (init_to_uniform) and near the simulating truth."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy
    from oracle import c_oracle

    data, true = simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, random_seed=0, **sim_kw)
    X, W, y = (data[k].astype(np.float32) for k in ("site_covs", "obs_covs", "obs"))
    T = data.get("session_duration")
    T = None if T is None else T.astype(np.float32)
    rng = np.random.default_rng(5)
    truth = np.concatenate([true["beta"][0], true["alpha"][0]])
    with bb.OccupancyLikelihood(model, X, W, y, T, max_chains=256, **model_kw) as lk:
        D = lk.theta_dim
        t0 = np.concatenate([truth, np.full(D - truth.size, np.log(0.1))])
        th = np.concatenate([rng.uniform(-2, 2, size=(128, D)), t0 + 1e-2 * rng.standard_normal((128, D))]).astype(np.float32)
        lp, gr = lk.logp_and_grad(th)
    idx = [0, 77, 127, 128, 200, 255]
    fn = c_oracle.occu_rn_logp_grad if model == "occu_rn" else c_oracle.occu_cop_logp_grad
    args = (X, W, y) if model == "occu_rn" else (X, W, y, T)
    ref_lp, ref_gr = fn(th[idx].astype(np.float64), *args, **okw)
    assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"{model} full size")


@pytest.mark.parametrize("name", ["occu_5x3", "occu_missing", "rn_5x3", "cop_missing_5x3", "nmix_missing_5x3",
                                  "nmix_default", "cs_missing_5x3", "cs_default"])
def test_strict_math_flag(name):
    """BL_FLAG_STRICT_MATH (libm expf/log1pf/IEEE division) and the default bounded-error SFU forms
    both meet the fp32 tolerance; their mutual difference is at fp32-rounding level."""
    g = load_golden(name)
    th = np.tile(g["thetas"], (20, 1))  # 140 chains -> the chain-parallel kernels where they exist
    ref_lp = np.tile(g["logp_f32"], 20)
    ref_gr = np.tile(g["grad_f32"], (20, 1))
    with _make(g, "float32", strict_math=True) as strict, _make(g, "float32") as fast:
        lp_s, gr_s = strict.logp_and_grad(th)
        lp_f, gr_f = fast.logp_and_grad(th)
    assert_close(lp_s, gr_s, ref_lp, ref_gr, 1e-5, f"{name}/strict")
    assert_close(lp_f, gr_f, ref_lp, ref_gr, 1e-5, f"{name}/sfu")


@pytest.mark.parametrize("ks,ko,K,fpc", [(1, 1, 100, False), (5, 3, 50, False), (5, 3, 30, True), (1, 1, 12, True),
                                          (0, 0, 20, False), (8, 4, 25, True), (3, 2, 40, False)])
def test_rn_chain_kernel_against_oracle(ks, ko, K, fpc):
    """The lane=chain Royle-Nichols kernel (C >= 32): ragged visits, NaN covariates, clamp-active
    states (r close to 1 makes k*log(1-r) cross log eps), optional false-positive constant."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(K + ks)
    S, J = 157, 11
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, 1, J, ko)) * 1.5
    y = (rng.uniform(size=(1, S, 1, J)) < 0.3).astype(float)
    lens = rng.integers(0, J + 1, size=S)
    y[0, :, 0][np.arange(J)[None, :] >= lens[:, None]] = np.nan
    W[rng.uniform(size=W.shape) < 0.03] = np.nan
    D = ks + ko + 2 + int(fpc)
    th = rng.uniform(-2, 2, size=(150, D))
    pr = orc.prepare(X, W, y)
    idx = [0, 1, 33, 149]
    ref_lp, ref_gr = orc.logp_grad("occu_rn", th[idx], pr, max_abundance=K, fp_constant=fpc)
    with bb.OccupancyLikelihood("occu_rn", X, W, y, max_abundance=K, false_positives_constant=fpc) as lk:
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"rn chain ks={ks} K={K} fpc={fpc}")
        lp_e, gr_e = lk.logp_and_grad(th[:8])  # same handle, site-parallel engine (C < 32)
        np.testing.assert_allclose(lp_e, lp[:8], rtol=5e-6)
        for n in (40, 128, 150):  # 128- and 256-thread chain blocks, partly idle warps
            lp_s, gr_s = lk.logp_and_grad(th[:n])
            np.testing.assert_allclose(lp_s, lp[:n], rtol=5e-6)
            np.testing.assert_allclose(gr_s, gr[:n], rtol=1e-4, atol=1e-5 * np.abs(gr).max())


@pytest.mark.parametrize("ks,ko,fpc,fpu", [(1, 1, True, False), (5, 3, True, True), (5, 3, False, False), (1, 1, False, True),
                                           (0, 2, True, False), (8, 4, True, True), (3, 1, False, False)])
def test_cop_chain_kernel_against_oracle(ks, ko, fpc, fpu):
    """The lane=chain count-detection kernel (C >= 32): ragged / missing visits, varying exposure, every
    false-positive configuration (without any, a unit that saw a count is occupied with certainty)."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(ks * 10 + ko + int(fpc) + 2 * int(fpu))
    S, J = 203, 12 if (ks == 5 and ko == 3) else 7
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, 1, J, ko)) * 0.7
    y = rng.poisson(1.5, size=(1, S, 1, J)).astype(float)
    y[0, rng.uniform(size=S) < 0.3] = 0.0  # sites without any count
    lens = rng.integers(0, J + 1, size=S)
    y[0, :, 0][np.arange(J)[None, :] >= lens[:, None]] = np.nan
    W[rng.uniform(size=W.shape) < 0.03] = np.nan
    T = rng.uniform(0.5, 9.0, size=(S, 1, J))
    D = ks + ko + 2 + int(fpc) + int(fpu)
    th = rng.uniform(-1.5, 1.5, size=(130, D))
    pr = orc.prepare(X, W, y, T)
    idx = [0, 1, 50, 129]
    ref_lp, ref_gr = orc.logp_grad("occu_cop", th[idx], pr, fp_constant=fpc, fp_unoccupied=fpu)
    with bb.OccupancyLikelihood("occu_cop", X, W, y, T, false_positives_constant=fpc,
                                false_positives_unoccupied=fpu) as lk:
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"cop chain ks={ks} fpc={fpc} fpu={fpu}")
        lp_e, gr_e = lk.logp_and_grad(th[:8])  # site-parallel engine on the same handle
        np.testing.assert_allclose(lp_e, lp[:8], rtol=5e-6)
        for n in (40, 100, 128):  # 128- and 256-thread chain blocks, partly idle warps
            lp_s, gr_s = lk.logp_and_grad(th[:n])
            np.testing.assert_allclose(lp_s, lp[:n], rtol=5e-6)
            np.testing.assert_allclose(gr_s, gr[:n], rtol=1e-4, atol=1e-5 * np.abs(gr).max())


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("name", ["occu_missing", "occu_fp_const", "rn_5x3", "cop_missing_5x3", "cop_both_fp",
                                  "nmix_missing_5x3", "cs_missing_5x3"])
def test_site_summary_streaming_kernel(name, dtype):
    """bl_site_summary (psi / occupancy probability / pointwise lppd and p_waic per unit, streamed over
    draws) against the oracle's per-site terms."""
    from oracle import occupancy as orc

    g = load_golden(name)
    d = g["data"]
    rng = np.random.default_rng(1)
    draws = g["thetas"][0] + 0.1 * rng.standard_normal((45, g["thetas"].shape[1]))
    pr = orc.prepare(d["site_covs"], d["obs_covs"], d["obs"], d.get("session_duration"),
                     dtype=np.float32 if dtype == "float32" else np.float64)
    ref = orc.site_summary(g["model"], draws.astype(np.float32 if dtype == "float32" else np.float64), pr,
                           dtype=np.float32 if dtype == "float32" else np.float64, **g["model_kwargs"])
    with _make(g, dtype, prior=False) as lk:
        out = lk.site_summary(draws)
    rn = g["model"] in ("occu_rn", "nmixture")
    tol = 2e-5 if dtype == "float32" else 1e-6  # outputs are float32 either way
    a1 = out["abundance_mean" if rn else "psi_mean"].ravel()
    a2 = out["abundance_posterior_mean" if rn else "occupancy_prob"].ravel()
    np.testing.assert_allclose(a1, ref["a1"], rtol=tol, atol=tol)
    np.testing.assert_allclose(a2, ref["a2"], rtol=tol, atol=tol)
    np.testing.assert_allclose(out["lppd"].ravel(), ref["lppd"], rtol=tol, atol=20 * tol)
    np.testing.assert_allclose(out["p_waic"].ravel(), ref["p_waic"], rtol=2e-3, atol=1e-5)
    assert abs(out["lppd_total"] - ref["lppd"].sum()) <= 1e-5 * abs(ref["lppd"].sum())


@pytest.mark.parametrize("ks,ko", [(0, 1), (3, 2), (4, 4), (8, 3), (7, 1), (2, 2)])
def test_chain_kernel_runtime_ks_variants(ks, ko):
    """lane=chain occu kernel with a runtime number of site covariates (any Ks <= 8, Ko in 1..4)."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(ks * 5 + ko)
    S, J = 301, 9
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, 1, J, ko))
    y = (rng.uniform(size=(1, S, 1, J)) < 0.35).astype(float)
    y[0, rng.uniform(size=S) < 0.4] = 0.0
    y[rng.uniform(size=y.shape) < 0.1] = np.nan
    if ks:
        X[5, 0] = np.nan
    D = ks + ko + 2
    th = rng.uniform(-2, 2, size=(140, D))
    pr = orc.prepare(X, W, y)
    idx = [0, 7, 139]
    ref_lp, ref_gr = orc.logp_grad("occu", th[idx], pr)
    with bb.OccupancyLikelihood("occu", X, W, y) as lk:
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"runtime-Ks chain ks={ks} ko={ko}")


@pytest.mark.parametrize("ks,ko", [(1, 1), (5, 3), (0, 2), (8, 4), (3, 1), (2, 3)])
def test_cs_chain_kernel_against_oracle(ks, ko):
    """The lane=chain continuous-score kernel (C >= 32): ragged / missing visits, NaN covariates, score
    parameters from far off (sigma = e^-1.5: |n1 - n0| in the thousands) to the generating values."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(ks * 7 + ko)
    S, J = 171, 9
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, 1, J, ko))
    occ = rng.uniform(size=(1, S, 1, 1)) < 0.5
    f = (rng.uniform(size=(1, S, 1, J)) < 0.4) & occ
    y = np.where(f, rng.normal(10, 5, size=f.shape), rng.normal(0, 10, size=f.shape))
    lens = rng.integers(0, J + 1, size=S)
    y[0, :, 0][np.arange(J)[None, :] >= lens[:, None]] = np.nan
    W[rng.uniform(size=W.shape) < 0.03] = np.nan
    D = ks + ko + 6
    th = rng.uniform(-1.5, 1.5, size=(300, D))
    th[::3, -4:] = np.array([0.0, np.log(10.0), np.log(10.0), np.log(5.0)]) + 0.1 * rng.standard_normal((100, 4))
    pr = orc.prepare(X, W, y)
    idx = [0, 1, 2, 50, 131, 299]
    ref_lp, ref_gr = orc.logp_grad("occu_cs", th[idx], pr)
    with bb.OccupancyLikelihood("occu_cs", X, W, y) as lk:
        lp, gr = lk.logp_and_grad(th)  # 300 chains: three 128-thread chunks of 100
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"cs chain ks={ks} ko={ko}")
        lp_e, gr_e = lk.logp_and_grad(th[:8])  # site-parallel engine on the same handle
        np.testing.assert_allclose(lp_e, lp[:8], rtol=5e-6)
        for n in (40, 128, 256):  # 128- and 256-thread chain blocks, partly idle warps
            lp_s, gr_s = lk.logp_and_grad(th[:n])
            np.testing.assert_allclose(lp_s, lp[:n], rtol=5e-6)
            np.testing.assert_allclose(gr_s, gr[:n], rtol=1e-4, atol=1e-5 * np.abs(gr).max())


@pytest.mark.parametrize("model,sim_kw,model_kw", [
    ("occu_rn", dict(n_sites=200_000, deployment_days_per_site=70), dict(max_abundance=50)),
    ("occu_cop", dict(n_sites=500_000, deployment_days_per_site=84, simulate_missing=True),
     dict(false_positives_constant=True)),
    ("occu_cs", dict(n_sites=500_000, deployment_days_per_site=70), {}),
    ("nmixture", dict(n_sites=200_000, deployment_days_per_site=70), dict(max_abundance=80)),
])
def test_full_size_site_split_additivity(model, sim_kw, model_kw):
    """BASELINE configs 3 / 4 (and the two siblings at the same scale) at FULL size, through a
    size-independent property: the log-likelihood and its gradient are sums over sites, so evaluating two
    site blocks on their own handles and adding must reproduce the whole (40 chains: the lane=chain kernels
    for occu_rn / occu_cop, the site-parallel engine for the siblings)."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy

    data, _ = simulate_occupancy(model, n_site_covs=5, n_obs_covs=3, random_seed=0, **sim_kw)
    X = data["site_covs"].astype(np.float32)
    W = data["obs_covs"].astype(np.float32)
    y = data["obs"].astype(np.float32)
    T = data.get("session_duration")
    T = None if T is None else T.astype(np.float32)
    S = X.shape[0]
    rng = np.random.default_rng(3)

    def make(sl):
        return bb.OccupancyLikelihood(model, X[sl], W[sl], y[:, sl], None if T is None else T[sl], prior=False,
                                      **model_kw)

    with make(slice(None)) as full:
        D = full.theta_dim
        th = rng.uniform(-1, 1, size=(40, D))
        if model == "occu_cs":
            th[:, -4:] = np.array([0.0, np.log(10.0), np.log(10.0), np.log(5.0)]) + 0.1 * rng.standard_normal((40, 4))
        lp, gr = full.logp_and_grad(th)
    h = S // 3
    with make(slice(0, h)) as a, make(slice(h, S)) as b:
        la, ga = a.logp_and_grad(th)
        lb, gb = b.logp_and_grad(th)
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(gr))
    np.testing.assert_allclose(la.astype(np.float64) + lb, lp, rtol=5e-6)
    np.testing.assert_allclose(ga.astype(np.float64) + gb, gr, rtol=1e-4, atol=2e-5 * np.abs(gr).max())


@pytest.mark.parametrize("name", ["occu_missing", "occu_5x3"])
@pytest.mark.parametrize("dtype,tol", [("float32", 2e-5), ("float64", 3e-7)])  # outputs are float32
def test_pointwise_loglik_against_the_reference_closed_form(name, dtype, tol):
    """bl_obs_loglik vs numbers produced by EXECUTING the reference's log_likelihood_manual / lppd_manual
    (biolith/evaluation/log_likelihood.py:55-98, lppd.py:64-106) on the deterministic sites of the executed occu body
    (tests/golden/manual_refbody.npz): per-observation log-mean-exp, its variance over draws, and the lppd total."""
    import os

    from conftest import GOLDEN_DIR

    g = load_golden(name)
    m = np.load(os.path.join(GOLDEN_DIR, "manual_refbody.npz"))
    llm = m[f"{name}__log_lik_manual"][:, 0]          # (draws, S, P, J), NaN where obs is NaN
    with _make(g, dtype, prior=False) as lk:
        out = lk.pointwise_loglik(g["thetas"])
        mask = lk.mask()
    n = llm.shape[0]
    mx = np.nanmax(llm, axis=0)
    ref_lppd = mx + np.log(np.exp(llm - mx).sum(axis=0) / n)
    ref_var = llm.var(axis=0, ddof=1)
    assert np.array_equal(np.isfinite(out["lppd"]), mask)
    np.testing.assert_allclose(out["lppd"][mask], ref_lppd[mask], rtol=tol, atol=tol)
    np.testing.assert_allclose(out["p_waic"][mask], ref_var[mask], rtol=200 * tol, atol=tol)
    assert abs(out["lppd_total"] - float(m[f"{name}__lppd_manual"])) <= 10 * tol * abs(float(m[f"{name}__lppd_manual"]))


@pytest.mark.parametrize("tag,model,extra,flags", [
    ("occu_sp2_fpc", "occu", "prob_fp_constant", dict(false_positives_constant=True)),
    ("occu_p2_sp3", "occu", None, {}),
    ("cop_sp2_fpu", "occu_cop", "rate_fp_unoccupied", dict(false_positives_unoccupied=True)),
])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_multi_species_against_the_executed_reference_body(tag, model, extra, flags, dtype):
    """n_species > 1 with a false-positive parameter shared across the species plate (occu.py:146-157 sampled outside
    occu.py:182): the composite handle (csrc/multi.cu) against the executed reference body -- log joint, every
    species' beta / alpha gradient, and the gradient of the shared extra (a sum over species)."""
    import os

    import biolith_b200 as bb
    from conftest import GOLDEN_DIR

    e = dict(np.load(os.path.join(GOLDEN_DIR, "extra_refbody.npz")))
    X, W, y = (e[f"{tag}__data__{k}"] for k in ("site_covs", "obs_covs", "obs"))
    T = e.get(f"{tag}__data__session_duration")
    tol = 1e-5 if dtype == "float32" else 1e-10
    n = e[f"{tag}__logp"].size
    th, gref = [], []
    for i in range(n):
        parts = [e[f"{tag}__param__beta"][i].ravel(), e[f"{tag}__param__alpha"][i].ravel()]
        gparts = [e[f"{tag}__grad__beta"][i].ravel(), e[f"{tag}__grad__alpha"][i].ravel()]
        if extra:
            parts.append(np.atleast_1d(e[f"{tag}__param__{extra}"][i]))
            gparts.append(np.atleast_1d(e[f"{tag}__grad__{extra}"][i]))
        th.append(np.concatenate(parts))
        gref.append(np.concatenate(gparts))
    th, gref = np.stack(th), np.stack(gref)
    # the fixtures were generated with float64 data and float64 clamp constants; fp32 runs see rounded data
    with bb.OccupancyLikelihood(model, X, W, y, T, dtype=dtype, prior=True, **flags) as lk:
        assert lk.n_species == y.shape[0] and lk.theta_dim == th.shape[1]
        lp, gr = lk.logp_and_grad(th)
        assert lk.mask().shape == y.shape
    assert_close(lp, gr, e[f"{tag}__logp"], gref, tol if dtype == "float64" else 3e-5, f"{tag}/{dtype}")


def test_fit_two_species_with_a_shared_false_positive_parameter():
    """fit() on two species that share prob_fp_constant: one joint NUTS run over the composite handle; sample sites
    have numpyro's plate layout (beta: (draws, n_species, Kb))."""
    import biolith_b200 as bb
    from biolith_b200.simulate import simulate_occupancy

    data, true = simulate_occupancy("occu", n_species=2, n_sites=300, deployment_days_per_site=70,
                                    prob_fp_constant=0.05, random_seed=1)
    res = bb.fit(bb.models.occu, data["site_covs"], data["obs_covs"], data["obs"], num_chains=8, num_warmup=300,
                 num_samples=200, false_positives_constant=True)
    smp = res.samples
    assert smp["cov_state_0"].shape == (8 * 200, 2) and smp["prob_fp_constant"].shape == (8 * 200,)
    assert smp["psi"].shape[-1] == 2
    assert abs(smp["prob_fp_constant"].mean() - 0.05) < 0.05
    np.testing.assert_allclose(smp["cov_state_0"].mean(axis=0), true["beta"][:, 0], atol=0.6)
