"""K1s (csrc/occu_small.cu): the lane = site occu kernel that serves batches of fewer than 32 chains -- what `fit`
with the reference's default num_chains = 5 evaluates (biolith/utils/fit.py:24).  Parity against the oracle on every
code path: register-decoded J = 8 shape, runtime-J quads, chain chunks on grid.y, the clamp fallback, padded tiles."""

import numpy as np
import pytest

from test_gpu_parity import assert_close

pytestmark = pytest.mark.gpu


def _data(rng, S, P, J, ks, ko, missing=True):
    X = rng.normal(size=(S, ks))
    W = rng.normal(size=(S, P, J, ko))
    y = (rng.uniform(size=(1, S, P, J)) < 0.35).astype(float)
    y[0, rng.uniform(size=S) < 0.4] = 0.0
    if missing:
        y[rng.uniform(size=y.shape) < 0.15] = np.nan
        W[rng.uniform(size=W.shape) < 0.03] = np.nan
        if ks:
            X[rng.uniform(size=X.shape) < 0.03] = np.nan
    return X.astype(np.float32), W.astype(np.float32), y.astype(np.float32)


@pytest.mark.parametrize("ks,ko,J,P", [(5, 3, 8, 1), (5, 3, 8, 2), (2, 2, 5, 1), (0, 1, 3, 2), (8, 4, 13, 1),
                                       (3, 1, 33, 1), (1, 1, 1, 1), (4, 2, 12, 1)])
def test_small_kernel_against_oracle(ks, ko, J, P):
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(100 * ks + 10 * ko + J)
    S = 2077  # not a multiple of 32: the last warp-tile is padded
    X, W, y = _data(rng, S, P, J, ks, ko)
    pr = orc.prepare(X, W, y)
    D = ks + ko + 2
    with bb.OccupancyLikelihood("occu", X, W, y) as lk:
        for C in (1, 2, 5, 7, 8, 9, 17, 31):
            assert lk.plan(C)["kernel"] == 7, "batches below 32 chains are meant to run on K1s"
            th = rng.uniform(-2, 2, size=(C, D)).astype(np.float32)
            ref_lp, ref_gr = orc.logp_grad("occu", th.astype(np.float64), pr)
            lp, gr = lk.logp_and_grad(th)
            assert_close(lp, gr, ref_lp, ref_gr, 1e-5, f"K1s ks={ks} ko={ko} J={J} P={P} C={C}")
        assert lk.plan(32)["kernel"] in (1, 5)
    with bb.OccupancyLikelihood("occu", X, W, y, prior=False) as lk:
        th = rng.uniform(-2, 2, size=(5, D)).astype(np.float32)
        ref_lp, ref_gr = orc.logp_grad("occu", th.astype(np.float64), pr, prior=False)
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp, gr, ref_lp, ref_gr, 1e-5, "K1s likelihood only")


@pytest.mark.parametrize("ks,ko,J", [(5, 3, 8), (2, 2, 6)])
def test_small_kernel_clamps(ks, ko, J):
    """Thetas large enough that visits sit at numpyro's clamps (p~ = tiny / 1 - eps): the per-lane fallback of K1s
    must reproduce clamp_probs' values and zeroed derivatives (biolith/models/occu.py:229-242 through numpyro)."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(7 + J)
    X, W, y = _data(rng, 999, 1, J, ks, ko)
    pr = orc.prepare(X, W, y)
    D = ks + ko + 2
    th = rng.uniform(-2, 2, size=(6, D))
    th[1] *= 6.0
    th[2] *= 15.0
    th[3, ks + 1:] *= 40.0   # detection side only
    th[4, :ks + 1] *= 40.0   # occupancy side only
    th[5] = 0.0
    th = th.astype(np.float32)
    ref_lp, ref_gr = orc.logp_grad("occu", th.astype(np.float64), pr)
    with bb.OccupancyLikelihood("occu", X, W, y) as lk:
        assert lk.plan(6)["kernel"] == 7
        lp, gr = lk.logp_and_grad(th)
    assert_close(lp, gr, ref_lp, ref_gr, 1e-5, "K1s at the clamps")


def test_strict_engine_switch(monkeypatch):
    """BL_STRICT_ENGINE=1 keeps strict math on the libm engine (the A/B reference of the STRICT instantiations)."""
    import biolith_b200 as bb

    rng = np.random.default_rng(5)
    X, W, y = _data(rng, 1500, 1, 8, 5, 3)
    th = rng.uniform(-2, 2, size=(40, 10)).astype(np.float32)
    with bb.OccupancyLikelihood("occu", X, W, y, strict_math=True) as lk:
        a = [lk.logp_and_grad(th), lk.logp_and_grad(th[:6])]
    monkeypatch.setenv("BL_STRICT_ENGINE", "1")
    with bb.OccupancyLikelihood("occu", X, W, y, strict_math=True) as lk:
        assert lk.plan(40)["kernel"] == 0 and lk.plan(6)["kernel"] == 0
        b = [lk.logp_and_grad(th), lk.logp_and_grad(th[:6])]
    for (lp, gr), (lp0, gr0) in zip(a, b):
        np.testing.assert_allclose(lp, lp0, rtol=2e-6)
        np.testing.assert_allclose(gr, gr0, rtol=2e-5, atol=2e-5 * np.abs(gr0).max())


def test_small_kernel_matches_engine_and_repeats(monkeypatch):
    """Same numbers (to fp32 rounding) as the site-parallel engine it replaces, bit-identical on repetition, and
    independent of how the chains are batched (a chain's result never depends on its neighbours)."""
    import biolith_b200 as bb

    rng = np.random.default_rng(11)
    X, W, y = _data(rng, 40_000, 1, 8, 5, 3)
    th = rng.uniform(-2, 2, size=(31, 10)).astype(np.float32)
    with bb.OccupancyLikelihood("occu", X, W, y) as lk:
        p = lk.plan(31)
        assert p["kernel"] == 7 and p["grid"][1] == 4  # 31 chains = 4 chunks of <= 8
        lp, gr = lk.logp_and_grad(th)
        lp2, gr2 = lk.logp_and_grad(th)
        assert np.array_equal(lp, lp2) and np.array_equal(gr, gr2), "not deterministic"
        lp5, gr5 = lk.logp_and_grad(th[:5])
        np.testing.assert_allclose(lp5, lp[:5], rtol=2e-7)
        np.testing.assert_allclose(gr5, gr[:5], rtol=1e-5, atol=1e-5 * np.abs(gr).max())
    monkeypatch.setenv("BL_SMALL_KERNEL", "0")
    with bb.OccupancyLikelihood("occu", X, W, y) as lk:
        assert lk.plan(31)["kernel"] == 0
        lp0, gr0 = lk.logp_and_grad(th)
    np.testing.assert_allclose(lp, lp0, rtol=1e-6)
    np.testing.assert_allclose(gr, gr0, rtol=1e-5, atol=1e-5 * np.abs(gr0).max())


@pytest.mark.parametrize("ks,ko,J", [(5, 3, 8), (2, 2, 5), (8, 4, 13), (1, 1, 3)])
def test_strict_math_runs_the_chain_kernel_with_libm(ks, ko, J):
    """BL_FLAG_STRICT_MATH (north_star's "fast-math-free expf / log1pf") on occu with >= 32 chains: K1d's STRICT
    instantiations (FMA-pipe exp2, libm log2f, IEEE division, the engine's libm clamp form in the fallback), not the
    2.2 x slower engine (which still serves strict math below 32 chains); same 1e-5 bar, including thetas at the clamps."""
    import biolith_b200 as bb
    from oracle import occupancy as orc

    rng = np.random.default_rng(31 * ks + J)
    X, W, y = _data(rng, 3001, 1, J, ks, ko)
    pr = orc.prepare(X, W, y)
    D = ks + ko + 2
    th = rng.uniform(-2, 2, size=(70, D))
    th[1] *= 6.0
    th[2] *= 15.0
    th = th.astype(np.float32)
    idx = [0, 1, 2, 33, 69]
    ref_lp, ref_gr = orc.logp_grad("occu", th[idx].astype(np.float64), pr)
    with bb.OccupancyLikelihood("occu", X, W, y, strict_math=True) as lk:
        assert lk.plan(70)["kernel"] == 5 and lk.plan(5)["kernel"] == 0  # K1d STRICT / the libm engine
        lp, gr = lk.logp_and_grad(th)
        assert_close(lp[idx], gr[idx], ref_lp, ref_gr, 1e-5, f"strict K1d ks={ks} ko={ko} J={J}")
        lp5, gr5 = lk.logp_and_grad(th[:5])  # the libm engine on the same chains
        np.testing.assert_allclose(lp5, lp[:5], rtol=2e-6)
        np.testing.assert_allclose(gr5, gr[:5], rtol=2e-5, atol=2e-5 * np.abs(gr[:5]).max())
