"""ctypes loader of oracle/liboccu_oracle.so (the C/OpenMP restatement).  TEST INFRASTRUCTURE ONLY."""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboccu_oracle.so")
_lib = None


def load(build_if_missing=True):
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            if not build_if_missing:
                raise FileNotFoundError(_LIB)
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB)
        _lib.oracle_occu_logp_grad.restype = C.c_int
        _lib.oracle_occu_logp_grad.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_occu_cop_logp_grad.restype = C.c_int
        _lib.oracle_occu_cop_logp_grad.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_occu_rn_logp_grad.restype = C.c_int
        _lib.oracle_occu_rn_logp_grad.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_nmixture_logp_grad.restype = C.c_int
        _lib.oracle_nmixture_logp_grad.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_occu_cs_logp_grad.restype = C.c_int
        _lib.oracle_occu_cs_logp_grad.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def max_threads():
    return load().oracle_max_threads()


def occu_logp_grad(theta, site_covs, obs_covs, obs, dtype=np.float32, prior=True, nthreads=0):
    """site_covs (S,Ks), obs_covs (S,P,J,Ko), obs (1,S,P,J) in the reference layout."""
    lib = load()
    X = np.ascontiguousarray(site_covs, dtype=dtype)
    W = np.ascontiguousarray(obs_covs, dtype=dtype)
    y = np.ascontiguousarray(obs, dtype=dtype)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    S, P, J, Ko = W.shape
    Ks = X.shape[1]
    assert y.shape == (1, S, P, J) and th.shape[1] == Ks + Ko + 2
    n = th.shape[0]
    logp = np.empty(n)
    grad = np.empty((n, th.shape[1]))
    rc = lib.oracle_occu_logp_grad(0 if dtype == np.float32 else 1, S, P, J, Ks, Ko, y.ctypes.data, X.ctypes.data,
                                   W.ctypes.data, th.ctypes.data, n, int(prior), int(nthreads), logp.ctypes.data,
                                   grad.ctypes.data)
    assert rc == 0
    return logp, grad


def occu_cop_logp_grad(theta, site_covs, obs_covs, obs, session_duration=None, dtype=np.float32, prior=True,
                       fp_constant=False, fp_unoccupied=False, nthreads=0):
    """occu_cop in double arithmetic with the clamp constants of ``dtype`` (the dtype the reference would run
    in; arrays are first rounded to it, as the reference's ingestion does).  Reference layout as above."""
    lib = load()
    X = np.ascontiguousarray(np.asarray(site_covs, dtype=dtype), dtype=np.float64)
    W = np.ascontiguousarray(np.asarray(obs_covs, dtype=dtype), dtype=np.float64)
    y = np.ascontiguousarray(np.asarray(obs, dtype=dtype), dtype=np.float64)
    T = None if session_duration is None else np.ascontiguousarray(session_duration, dtype=np.float64)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    S, P, J, Ko = W.shape
    Ks = X.shape[1]
    D = Ks + Ko + 2 + int(bool(fp_constant)) + int(bool(fp_unoccupied))
    assert y.shape == (1, S, P, J) and th.shape[1] == D and (T is None or T.shape == (S, P, J))
    n = th.shape[0]
    logp = np.empty(n)
    grad = np.empty((n, D))
    rc = lib.oracle_occu_cop_logp_grad(int(dtype == np.float32), S, P, J, Ks, Ko, y.ctypes.data, X.ctypes.data,
                                       W.ctypes.data, None if T is None else T.ctypes.data, th.ctypes.data, n,
                                       int(bool(fp_constant)), int(bool(fp_unoccupied)), int(prior), int(nthreads),
                                       logp.ctypes.data, grad.ctypes.data)
    assert rc == 0
    return logp, grad


def occu_rn_logp_grad(theta, site_covs, obs_covs, obs, max_abundance=100, dtype=np.float32, prior=True,
                      fp_constant=False, nthreads=0):
    """occu_rn in double arithmetic with the clamp constants of ``dtype``.  Reference layout as above."""
    lib = load()
    X = np.ascontiguousarray(np.asarray(site_covs, dtype=dtype), dtype=np.float64)
    W = np.ascontiguousarray(np.asarray(obs_covs, dtype=dtype), dtype=np.float64)
    y = np.ascontiguousarray(np.asarray(obs, dtype=dtype), dtype=np.float64)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    S, P, J, Ko = W.shape
    Ks = X.shape[1]
    D = Ks + Ko + 2 + int(bool(fp_constant))
    assert y.shape == (1, S, P, J) and th.shape[1] == D
    n = th.shape[0]
    logp = np.empty(n)
    grad = np.empty((n, D))
    rc = lib.oracle_occu_rn_logp_grad(int(dtype == np.float32), S, P, J, Ks, Ko, int(max_abundance), y.ctypes.data,
                                      X.ctypes.data, W.ctypes.data, th.ctypes.data, n, int(bool(fp_constant)),
                                      int(prior), int(nthreads), logp.ctypes.data, grad.ctypes.data)
    assert rc == 0
    return logp, grad


def nmixture_logp_grad(theta, site_covs, obs_covs, obs, max_abundance=100, dtype=np.float32, prior=True, nthreads=0):
    """nmixture in double arithmetic (arrays first rounded to ``dtype``).  Reference layout as above."""
    lib = load()
    X = np.ascontiguousarray(np.asarray(site_covs, dtype=dtype), dtype=np.float64)
    W = np.ascontiguousarray(np.asarray(obs_covs, dtype=dtype), dtype=np.float64)
    y = np.ascontiguousarray(np.asarray(obs, dtype=dtype), dtype=np.float64)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    S, P, J, Ko = W.shape
    Ks = X.shape[1]
    assert y.shape == (1, S, P, J) and th.shape[1] == Ks + Ko + 2
    n = th.shape[0]
    logp = np.empty(n)
    grad = np.empty((n, th.shape[1]))
    rc = lib.oracle_nmixture_logp_grad(int(dtype == np.float32), S, P, J, Ks, Ko, int(max_abundance), y.ctypes.data,
                                       X.ctypes.data, W.ctypes.data, th.ctypes.data, n, int(prior), int(nthreads),
                                       logp.ctypes.data, grad.ctypes.data)
    assert rc == 0
    return logp, grad


def occu_cs_logp_grad(theta, site_covs, obs_covs, obs, dtype=np.float32, prior=True, prior_mu_scale=10.0,
                      prior_sigma=(5.0, 1.0), nthreads=0):
    """occu_cs in double arithmetic with the clamp constants of ``dtype``.  Reference layout as above."""
    lib = load()
    X = np.ascontiguousarray(np.asarray(site_covs, dtype=dtype), dtype=np.float64)
    W = np.ascontiguousarray(np.asarray(obs_covs, dtype=dtype), dtype=np.float64)
    y = np.ascontiguousarray(np.asarray(obs, dtype=dtype), dtype=np.float64)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    S, P, J, Ko = W.shape
    Ks = X.shape[1]
    assert y.shape == (1, S, P, J) and th.shape[1] == Ks + Ko + 6
    n = th.shape[0]
    logp = np.empty(n)
    grad = np.empty((n, th.shape[1]))
    rc = lib.oracle_occu_cs_logp_grad(int(dtype == np.float32), S, P, J, Ks, Ko, y.ctypes.data, X.ctypes.data,
                                      W.ctypes.data, th.ctypes.data, n, int(prior), float(prior_mu_scale),
                                      float(prior_sigma[0]), float(prior_sigma[1]), int(nthreads), logp.ctypes.data,
                                      grad.ctypes.data)
    assert rc == 0
    return logp, grad
