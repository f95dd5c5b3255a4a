"""A minimal stand-in for the few jax entry points biolith_b200/jax_ffi.py uses, so that its registration,
custom_vjp, custom_vmap folding and drop-in models RUN on the GPU box (jax itself is not installable there).

TEST DOUBLE, not a product path.  Arrays are host numpy arrays; `jax.ffi.ffi_call` does what XLA's legacy
custom-call thunk does with a registered target: device buffers for operands and results, a NON-default stream,
`fn(stream, void** buffers, opaque, opaque_len, status)`, then a failure check.  numpyro / jax.numpy / jax.nn come
from oracle/refshim.py (the same stand-ins that execute the reference bodies).
"""

import ctypes as C
import os
import subprocess
import sys
import tempfile
import types

import numpy as np

_TARGETS = {}
_STUB = {}


class XlaRuntimeError(RuntimeError):
    pass


def status_stub():
    """A process-global `XlaCustomCallStatusSetFailure` (what the XLA runtime exports); records the message."""
    if not _STUB:
        src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "xla_status_stub.c")
        out = os.path.join(tempfile.mkdtemp(prefix="bl_xla_stub_"), "libxla_status_stub.so")
        subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-o", out, src], check=True)
        lib = C.CDLL(out, mode=C.RTLD_GLOBAL)
        lib.bl_test_status_message.restype = C.c_char_p
        lib.bl_test_status_reset.restype = None
        _STUB["lib"] = lib
    return _STUB["lib"]


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)


def register_ffi_target(name, capsule, platform="cpu", api_version=1):
    assert platform == "CUDA" and api_version == 0, "the library exports the legacy custom-call ABI"
    _TARGETS[name] = capsule


def pycapsule(fn):
    return fn  # a ctypes function pointer; the real jax wraps its address in a PyCapsule


def ffi_call(target, out_types, custom_call_api_version=None, legacy_backend_config=None, vmap_method=None, **kw):
    assert custom_call_api_version == 2, "legacy status-returning ABI"
    fn = _TARGETS[target]
    opaque = bytes(legacy_backend_config)

    def call(*operands):
        from biolith_b200 import _lib
        from biolith_b200.likelihood import DeviceBuffer

        lib = _lib.load()
        stub = status_stub()
        stub.bl_test_status_reset()
        stream = C.c_void_p()
        _lib.check(lib.bl_stream_create(0, C.byref(stream)), "bl_stream_create")
        bufs = []
        try:
            for a in operands:
                a = np.ascontiguousarray(a)
                b = DeviceBuffer(max(a.nbytes, 16))
                b.upload(a, stream)
                bufs.append(b)
            outs = []
            for t in out_types:
                n = int(np.prod(t.shape)) * t.dtype.itemsize
                outs.append(DeviceBuffer(max(n, 16)))
            ptrs = (C.c_void_p * (len(bufs) + len(outs)))(*[b.ptr.value for b in bufs + outs])
            status = C.c_void_p(0xB10)  # opaque token, only handed back to the failure callback
            fn(stream, ptrs, opaque, len(opaque), status)
            _lib.check(lib.bl_stream_sync(stream), "bl_stream_sync")
            msg = stub.bl_test_status_message()
            if msg:
                raise XlaRuntimeError(msg.decode())
            return tuple(o.download(t.shape, t.dtype, stream) for o, t in zip(outs, out_types))
        finally:
            for b in bufs + (outs if "outs" in locals() else []):
                b.free()
            lib.bl_stream_destroy(stream)

    return call


class custom_vjp:  # noqa: N801
    def __init__(self, f):
        self.f = f

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *a):
        return self.f(*a)


def vjp(f, *primals):
    assert isinstance(f, custom_vjp)
    out, res = f.fwd(*primals)
    return out, lambda ct: f.bwd(res, ct)


class custom_vmap:  # noqa: N801
    def __init__(self, f):
        self.f = f
        self.rule = None

    def def_vmap(self, rule):
        self.rule = rule
        return rule

    def __call__(self, *a):
        return self.f(*a)


def vmap(f):
    assert isinstance(f, custom_vmap) and f.rule is not None

    def mapped(*xs):
        out, _ = f.rule(xs[0].shape[0], [True] * len(xs), *xs)
        return out

    return mapped


def install():
    """refshim's numpyro / jax.numpy / jax.nn + the ffi / custom_vjp / custom_vmap doubles above."""
    from oracle import refshim

    mods = refshim.install()
    jax = mods["jax"]
    jax.ffi = types.SimpleNamespace(register_ffi_target=register_ffi_target, pycapsule=pycapsule, ffi_call=ffi_call)
    jax.ShapeDtypeStruct = ShapeDtypeStruct
    jax.custom_vjp = custom_vjp
    jax.vjp = vjp
    jax.vmap = vmap
    jax.custom_batching = types.SimpleNamespace(custom_vmap=custom_vmap)
    with np.errstate(divide="ignore"):
        mods["jax.scipy.special"].logit = lambda p: np.log(p) - np.log1p(-p)
    # the reference's regressor (biolith/regression/linear.py:16-66) is absent on the GPU box: a double with the
    # same two methods stands in when the real package cannot be imported
    try:
        import biolith.regression  # noqa: F401
    except Exception:
        numpyro = mods["numpyro"]

        class LinearRegression:
            def __init__(self, name, n_covs, prior=None):
                self.coef = numpyro.sample(name, (prior or numpyro.distributions.Normal()).expand([n_covs + 1]).to_event(1))

            def __call__(self, covs):
                return covs @ self.coef[..., 1:].T + self.coef[..., 0]

        pkg = types.ModuleType("biolith")
        pkg.__path__ = []
        reg = types.ModuleType("biolith.regression")
        reg.LinearRegression = LinearRegression
        pkg.regression = reg
        sys.modules["biolith"], sys.modules["biolith.regression"] = pkg, reg
    return mods


def uninstall():
    from oracle import refshim

    for k in ("biolith", "biolith.regression"):
        m = sys.modules.get(k)
        if m is not None and not getattr(m, "__file__", None):
            del sys.modules[k]
    refshim.uninstall()
