"""The jax.ffi boundary north_star names, exercised on the GPU without jax:

(a) `bl_xla_eval` called exactly as XLA's legacy GPU custom-call thunk calls it -- non-default stream,
    `void* buffers[3]` = {theta, logp, grad} device pointers, `bl_xla_opaque` bytes, a status token and the
    process-exported `XlaCustomCallStatusSetFailure` -- bitwise equal to `bl_eval`, failure path observable;
(b) biolith_b200/jax_ffi.py (registration, custom_vjp fwd/bwd, custom_vmap chain folding, drop-in models with the
    reference's sample sites and deterministic sites) run against tests/fakejax.py + oracle/refshim.py and
    compared with the executed reference bodies (tests/golden/*_refbody.npz).
"""

import ctypes as C
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

pytestmark = pytest.mark.gpu


def _lk(g, **kw):
    import biolith_b200 as bb

    d, mk = g["data"], g["model_kwargs"]
    return bb.OccupancyLikelihood(
        g["model"], d["site_covs"], d["obs_covs"], d["obs"], d.get("session_duration"),
        false_positives_constant=mk.get("fp_constant", False), false_positives_unoccupied=mk.get("fp_unoccupied", False),
        max_abundance=mk.get("max_abundance", 100), **kw)


def _xla_call(lk, theta, opaque=None, n_chains=None, null_buffers=False):
    """What XLA does with a registered legacy custom call.  Returns (logp, grad, failure message or None)."""
    import fakejax
    from biolith_b200 import _lib
    from biolith_b200.likelihood import DeviceBuffer

    lib = _lib.load()
    stub = fakejax.status_stub()
    stub.bl_test_status_reset()
    th = np.ascontiguousarray(theta, lk.np_dtype)
    n, D = th.shape
    stream = C.c_void_p()
    _lib.check(lib.bl_stream_create(0, C.byref(stream)), "bl_stream_create")
    assert stream.value, "XLA launches custom calls on its own (non-default) stream"
    d_th, d_lp, d_gr = DeviceBuffer(th.nbytes), DeviceBuffer(n * th.itemsize), DeviceBuffer(th.nbytes)
    d_th.upload(th, stream)
    d_lp.upload(np.full(n, np.nan, lk.np_dtype), stream)
    d_gr.upload(np.full((n, D), np.nan, lk.np_dtype), stream)
    bufs = (C.c_void_p * 3)(d_th.ptr.value, d_lp.ptr.value, d_gr.ptr.value)
    if opaque is None:
        opaque = bytes(_lib.bl_xla_opaque(dataset=lk.handle, n_chains=n if n_chains is None else n_chains, reserved=0))
    token = C.c_void_p(0x5a5a)
    lib.bl_xla_eval(stream, None if null_buffers else bufs, opaque, len(opaque), token)
    _lib.check(lib.bl_stream_sync(stream), "bl_stream_sync")
    msg = stub.bl_test_status_message()
    lp, gr = d_lp.download((n,), lk.np_dtype, stream), d_gr.download((n, D), lk.np_dtype, stream)
    for b in (d_th, d_lp, d_gr):
        b.free()
    lib.bl_stream_destroy(stream)
    return lp, gr, (msg.decode() if msg else None)


@pytest.mark.parametrize("name,dtype", [("occu_5x3", "float32"), ("occu_5x3", "float64"), ("rn_5x3", "float32"),
                                        ("cop_missing_5x3", "float32"), ("occu_fp_const", "float32")])
def test_xla_custom_call_equals_bl_eval_bitwise(name, dtype):
    g = load_golden(name)
    rng = np.random.default_rng(5)
    th = np.concatenate([g["thetas"], rng.uniform(-2, 2, (70, g["thetas"].shape[1]))])  # 77 chains: chain kernel too
    with _lk(g, dtype=dtype, prior=False) as lk:
        lp0, gr0 = lk.logp_and_grad(th)
        lp1, gr1, msg = _xla_call(lk, th)
        assert msg is None
        assert np.array_equal(lp0, lp1) and np.array_equal(gr0, gr1)
        mode = "f32" if dtype == "float32" else "f64"
        r = np.load(os.path.join(GOLDEN_DIR, name + "_refbody.npz"))
        tol = 1e-5 if dtype == "float32" else (1e-3 if g["model"] == "occu_rn" else 1e-10)
        k = len(g["thetas"])
        assert np.max(np.abs(lp1[:k] - r[f"ref_loglik_{mode}"]) / np.abs(r[f"ref_loglik_{mode}"])) < tol


def test_xla_custom_call_failure_path():
    g = load_golden("occu_default")
    with _lk(g) as lk:
        th = g["thetas"]
        # truncated opaque, null dataset, zero chains, null buffers: reported through XlaCustomCallStatusSetFailure,
        # outputs untouched (still the NaN fill), nothing launched
        from biolith_b200 import _lib

        n0 = _lib.load().bl_launch_count()
        for kw in (dict(opaque=b"\x00" * 8), dict(opaque=bytes(_lib.bl_xla_opaque(dataset=0, n_chains=7, reserved=0))),
                   dict(n_chains=0), dict(null_buffers=True)):
            lp, gr, msg = _xla_call(lk, th, **kw)
            assert msg and msg.startswith("biolith_b200"), (kw, msg)
            assert np.isnan(lp).all() and np.isnan(gr).all()
        assert _lib.load().bl_launch_count() == n0
        lp, gr, msg = _xla_call(lk, th)  # and the handle still works afterwards
        assert msg is None and np.isfinite(lp).all()


@pytest.fixture
def jax_ffi_module():
    import fakejax

    fakejax.install()
    import biolith_b200.jax_ffi as jf

    jf = importlib.reload(jf)
    assert jf.HAVE_JAX
    yield jf
    jf.close_datasets()
    fakejax.uninstall()
    importlib.reload(jf)


def test_make_loglik_vjp_and_vmap_fold(jax_ffi_module):
    import sys

    jf = jax_ffi_module
    jax = sys.modules["jax"]
    from biolith_b200 import _lib

    g = load_golden("occu_5x3")
    r = np.load(os.path.join(GOLDEN_DIR, "occu_5x3_refbody.npz"))
    with _lk(g, prior=False) as lk:
        loglik = jf.make_loglik(lk)
        th = g["thetas"].astype(np.float32)
        lp = loglik(th)                                       # primal: (C, D) -> (C,)
        np.testing.assert_allclose(lp, r["ref_loglik_f32"], rtol=1e-5)
        out, pull = jax.vjp(loglik, th)                       # fwd keeps the gradient, bwd scales it
        ct = np.linspace(0.5, 2.0, len(th)).astype(np.float32)
        (gth,) = pull(ct)
        np.testing.assert_allclose(out, lp, rtol=0)
        scale = np.abs(r["ref_gradlik_f32"]).max(axis=1, keepdims=True)
        assert np.max(np.abs(gth / ct[:, None] - r["ref_gradlik_f32"]) / scale) < 1e-5
        one = loglik(th[3])                                   # a single chain (what one numpyro chain calls)
        assert np.shape(one) == () and abs(float(one) - lp[3]) <= 2e-6 * abs(lp[3])
        # chain_method="vectorized": vmap over a leading axis folds into ONE launch of A*n chains
        th3 = np.stack([th[:6], th[:6][::-1], th[:6] + 0.01])  # (A=3, n=6, D)
        n0 = _lib.load().bl_launch_count()
        lp3, g3 = jax.vmap(loglik.batched_call)(th3)
        assert _lib.load().bl_launch_count() - n0 <= 2, "the chain axis must fold into one batched evaluation"
        assert lp3.shape == (3, 6) and g3.shape == (3, 6, th.shape[1])
        np.testing.assert_allclose(lp3[0], lp[:6], rtol=2e-6)
        np.testing.assert_allclose(lp3[1], lp[:6][::-1], rtol=2e-6)


@pytest.mark.parametrize("name", ["occu_5x3", "occu_missing", "occu_fp_const", "rn_5x3", "cop_missing_5x3",
                                  "nmix_missing_5x3", "cs_missing_5x3"])
def test_drop_in_model_equals_reference_body(jax_ffi_module, name):
    """The drop-in numpyro model (reference sample sites + numpyro.factor(kernel)) traced under the numpyro
    stand-in: log joint == the executed reference body, deterministic sites have the reference's shapes."""
    from oracle import refshim

    jf = jax_ffi_module
    g = load_golden(name)
    r = np.load(os.path.join(GOLDEN_DIR, name + "_refbody.npz"))
    d, mk = g["data"], g["model_kwargs"]
    model = getattr(jf, g["model"])
    kw = refshim.model_kwargs_for(g["model"], fp_constant=mk.get("fp_constant", False),
                                  fp_unoccupied=mk.get("fp_unoccupied", False), max_abundance=mk.get("max_abundance"))
    X32 = {k: np.asarray(v, np.float32) for k, v in d.items()}
    Ks, Ko = d["site_covs"].shape[1], d["obs_covs"].shape[3]
    with refshim.precision(np.float64, np.float32):
        for i in (0, 3, 6):
            params = refshim.theta_to_params(g["model"], g["thetas"][i], Ks, Ko, fp_constant=mk.get("fp_constant", False),
                                             fp_unoccupied=mk.get("fp_unoccupied", False))
            tr = refshim.trace(model, params, **X32, **kw)
            lj = float(tr.log_density())
            assert abs(lj - r["ref_logp_f32"][i]) <= 1e-5 * abs(r["ref_logp_f32"][i]), (name, i, lj)
            S, P, J = d["obs_covs"].shape[:3]
            occ = "abundance" if g["model"] in ("occu_rn", "nmixture") else "psi"
            assert tr.deterministic[occ].shape == (S, 1)
            if g["model"] not in ("occu_cs",):
                det = "rate_detection" if g["model"] == "occu_cop" else "prob_detection"
                assert tr.deterministic[det].shape == (J, P, S, 1)
    if name == "occu_missing":
        m = np.load(os.path.join(GOLDEN_DIR, "manual_refbody.npz"))
        with refshim.precision(np.float64):
            tr = refshim.trace(model, refshim.theta_to_params("occu", g["thetas"][2], Ks, Ko), **d)
        np.testing.assert_allclose(tr.deterministic["psi"], m["occu_missing__psi"][2], rtol=1e-12)
        np.testing.assert_allclose(tr.deterministic["prob_detection"], m["occu_missing__prob_detection"][2], rtol=1e-12)


def test_dataset_cache_is_keyed_by_content_and_bounded(jax_ffi_module):
    jf = jax_ffi_module
    g = load_golden("occu_default")
    d = g["data"]
    a = jf.dataset_handle("occu", d["site_covs"], d["obs_covs"], d["obs"], None, False, False, 100)
    b = jf.dataset_handle("occu", d["site_covs"].copy(), d["obs_covs"].copy(), d["obs"].copy(), None, False, False, 100)
    assert a is b, "equal content -> same packed dataset, whatever the object identity"
    y2 = d["obs"].copy()
    y2[0, 0, 0, 0] = 1 - np.nan_to_num(y2[0, 0, 0, 0])
    c = jf.dataset_handle("occu", d["site_covs"], d["obs_covs"], y2, None, False, False, 100)
    assert c is not a, "one changed observation -> a different dataset"
    for k in range(jf.MAX_CACHED_DATASETS + 1):
        yk = d["obs"].copy()
        yk[0, 1 + k, 0, :] = np.nan
        jf.dataset_handle("occu", d["site_covs"], d["obs_covs"], yk, None, False, False, 100)
    assert len(jf._handles) == jf.MAX_CACHED_DATASETS
    with pytest.raises(Exception):
        a.handle  # evicted handles were closed (device memory released)
    jf.close_datasets()
    assert not jf._handles
