"""Posterior-mode search on the device kernels (Adam on a small batch of starts, host-side numpy).

Not part of the reference (biolith/utils/fit.py:93 always starts NUTS from ``init_to_uniform``); it
exists because with 10^7 sites the posterior is ~10^-4 wide and a sampler started from U(-2, 2) with
numpyro's defaults (step size 1, tree depth 10) needs thousands of warm-up leapfrogs just to *reach*
the typical set.  ``fit(..., init_strategy="map")`` starts every chain at the mode found here plus a
small jitter; the target distribution is unchanged.  Works unchanged on a site-sharded handle (every
rank computes the same numbers, so every rank takes the same steps).
"""

from __future__ import annotations

import numpy as np


def find_map(likelihood, n_starts: int = 8, iters: int = 400, lr0: float = 0.1, lr1: float = 1e-3,
             seed: int = 0, init_radius: float = 2.0, verbose: bool = False):
    """Returns (theta_map (D,), logp_map, info).  The handle must include the priors."""
    D = likelihood.theta_dim
    rng = np.random.default_rng(seed)
    th = rng.uniform(-init_radius, init_radius, size=(n_starts, D))
    th[0] = 0.0
    m = np.zeros_like(th)
    v = np.zeros_like(th)
    b1, b2, eps = 0.9, 0.999, 1e-12
    best_lp = np.full(n_starts, -np.inf)
    best_th = th.copy()
    for t in range(1, iters + 1):
        lp, g = likelihood.logp_and_grad(th)
        lp, g = np.asarray(lp, np.float64), np.asarray(g, np.float64)
        ok = np.isfinite(lp) & np.all(np.isfinite(g), axis=1)
        better = ok & (lp > best_lp)
        best_lp = np.where(better, lp, best_lp)
        best_th[better] = th[better]
        g = np.where(ok[:, None], g, 0.0)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        lr = lr0 * (lr1 / lr0) ** ((t - 1) / max(iters - 1, 1))  # geometric decay
        th = th + lr * (m / (1 - b1**t)) / (np.sqrt(v / (1 - b2**t)) + eps)
        if verbose and (t % 50 == 0 or t == 1):
            print(f"[find_map] iter {t:4d} lr {lr:.2e} best logp {best_lp.max():.3f}")
    lp, g = likelihood.logp_and_grad(best_th)
    i = int(np.argmax(np.where(np.isfinite(lp), lp, -np.inf)))
    info = dict(logp_all=np.asarray(lp, np.float64), grad_inf_norm=float(np.abs(g[i]).max()), evals=(iters + 1) * n_starts)
    return best_th[i].astype(np.float64), float(lp[i]), info


def init_around(theta_map, num_chains: int, jitter: float = 1e-4, seed: int = 0):
    rng = np.random.default_rng(seed)
    return theta_map[None, :] + jitter * rng.standard_normal((num_chains, theta_map.size))
