"""Posterior diagnostics with the definitions of ``numpyro.diagnostics``.

The reference reports ESS / r-hat through ``numpyro.diagnostics.summary``
(biolith/evaluation/diagnostics.py:23-32).  numpyro is not vendored in the reference tree, so its
published estimators are restated here (numpyro/diagnostics.py: ``autocorrelation``,
``autocovariance``, ``effective_sample_size`` (Geyer initial monotone sequence over chains, as in
Stan), ``gelman_rubin``, ``split_gelman_rubin``, ``hpdi``, ``summary``).  Pure numpy, host side;
inputs are ``(num_chains, num_draws, ...)`` arrays.
"""

from __future__ import annotations

from collections import OrderedDict

import numpy as np


def _fft_next_fast_len(n: int) -> int:
    # smallest 2^a 3^b 5^c >= n
    while True:
        m = n
        for f in (2, 3, 5):
            while m % f == 0:
                m //= f
        if m == 1:
            return n
        n += 1


def autocorrelation(x, axis=0, bias=True):
    """Autocorrelation along ``axis`` via FFT, normalised at lag 0.  ``bias=True`` (numpyro >= 0.18's default,
    the version the reference pins: pyproject.toml:26) keeps the 1/N estimator; ``bias=False`` divides lag k by
    N - k (the older default)."""
    x = np.asarray(x, dtype=np.float64)
    N = x.shape[axis]
    M = _fft_next_fast_len(N)
    M2 = 2 * M
    x = np.swapaxes(x, axis, -1)
    centered = x - x.mean(axis=-1, keepdims=True)
    freq = np.fft.rfft(centered, n=M2, axis=-1)
    power = freq.real**2 + freq.imag**2
    ac = np.fft.irfft(power, n=M2, axis=-1)[..., :N]
    if not bias:
        ac = ac / np.arange(N, 0.0, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ac = ac / ac[..., :1]
    return np.swapaxes(ac, axis, -1)


def autocovariance(x, axis=0, bias=True):
    x = np.asarray(x, dtype=np.float64)
    return autocorrelation(x, axis, bias) * x.var(axis=axis, keepdims=True)


def _chain_variance_stats(x):
    # x: (C, N, ...)
    C, N = x.shape[0], x.shape[1]
    if N == 1:
        var_within = np.zeros(x.shape[2:])
    else:
        var_within = x.var(axis=1, ddof=1).mean(axis=0)
    var_estimator = var_within * (N - 1) / N
    if C > 1:
        var_estimator = var_estimator + x.mean(axis=1).var(axis=0, ddof=1)
    return var_within, var_estimator


def gelman_rubin(x):
    x = np.asarray(x, dtype=np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 2
    var_within, var_estimator = _chain_variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sqrt(var_estimator / var_within)


def split_gelman_rubin(x):
    x = np.asarray(x, dtype=np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 4
    N_half = x.shape[1] // 2
    new = np.concatenate([x[:, :N_half], x[:, -N_half:]], axis=0)
    return gelman_rubin(new)


def effective_sample_size(x, bias=True):
    """numpyro.diagnostics.effective_sample_size: x (num_chains, num_draws, ...) -> n_eff (...)."""
    x = np.asarray(x, dtype=np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 2
    C, N = x.shape[0], x.shape[1]
    gamma_k_c = autocovariance(x, axis=1, bias=bias)  # (C, N, ...)
    var_within, var_estimator = _chain_variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        rho_k = 1.0 - (var_within - gamma_k_c.mean(axis=0)) / var_estimator
    rho_k[0] = 1.0
    # pair sums, initial positive + initial monotone sequence
    Rho_k = rho_k[:-1:2, ...] + rho_k[1::2, ...]
    Rho_init = Rho_k[:1]
    Rho_pos = np.clip(Rho_k[1:, ...], 0.0, None)
    Rho_mono = np.minimum.accumulate(Rho_pos, axis=0)
    Rho_k = np.concatenate([Rho_init, Rho_mono], axis=0)
    tau = -1.0 + 2.0 * Rho_k.sum(axis=0)
    return C * N / tau


def hpdi(x, prob=0.90, axis=0):
    x = np.swapaxes(np.asarray(x, dtype=np.float64), axis, 0)
    sx = np.sort(x, axis=0)
    mass = sx.shape[0]
    idx_len = int(prob * mass)
    intervals_left = sx[: (mass - idx_len)]
    intervals_right = sx[idx_len:]
    idx_start = (intervals_right - intervals_left).argmin(axis=0)
    idx = np.expand_dims(idx_start, 0)
    lo = np.take_along_axis(sx, idx, axis=0)
    hi = np.take_along_axis(sx, idx + idx_len, axis=0)
    return np.swapaxes(np.concatenate([lo, hi], axis=0), axis, 0)


def summary(samples, prob=0.90, group_by_chain=True):
    """dict name -> {mean, std, median, 5.0%, 95.0%, n_eff, r_hat}; arrays are (chains, draws, ...)."""
    if not isinstance(samples, dict):
        samples = {"Param:0": samples}
    out = {}
    for name, value in samples.items():
        value = np.asarray(value, dtype=np.float64)
        if not group_by_chain:
            value = value[None, ...]
        flat = value.reshape((-1,) + value.shape[2:])
        lo, hi = hpdi(flat, prob=prob)
        d = OrderedDict(
            mean=flat.mean(axis=0), std=flat.std(axis=0, ddof=1), median=np.median(flat, axis=0))
        d[f"{50 * (1 - prob):.1f}%"] = lo
        d[f"{50 * (1 + prob):.1f}%"] = hi
        d["n_eff"] = effective_sample_size(value)
        d["r_hat"] = split_gelman_rubin(value) if value.shape[1] >= 4 else gelman_rubin(value)
        out[name] = d
    return out


def mcse_mean(x):
    """Monte Carlo standard error of the posterior mean: sd / sqrt(n_eff).  x (C, N, ...)."""
    x = np.asarray(x, dtype=np.float64)
    flat = x.reshape((-1,) + x.shape[2:])
    return flat.std(axis=0, ddof=1) / np.sqrt(effective_sample_size(x))
