"""CPU oracle for the occupancy log-density + gradient hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``biolith_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
timed CPU baseline, never as the product path.

PARITY STATUS: **unpinned against numpyro**.  The reference (timmh/biolith) only
*declares* the probabilistic program; the arithmetic of the path (enumeration
sum-product, clamped Bernoulli / Poisson / Categorical log-probs, MaskedDistribution,
reverse-mode AD) lives in numpyro>=0.18 / funsor>=0.4.5 / jax>=0.5 (lower bounds in
/root/reference/pyproject.toml:22-28, no lock file), none of which is installed or
installable here.  This file therefore restates (1) the reference's own model bodies
operation by operation and (2) the *published* numpyro semantics for the distributions
they call.  It is pinned three ways instead:
  * the data side is pinned to the reference's own ``simulate*()`` generators
    (tests/golden/*.npz, produced by tests/golden/make_golden.py importing
    /root/reference with jax/numpyro stubbed);
  * the op-by-op "enumerated" restatement and the closed-form restatement are
    independent derivations and must agree to ~1e-12;
  * closed-form gradients are checked against central finite differences of the
    enumerated restatement.

Two independent restatements are provided per model:

``*_log_joint_enumerated``  follows the reference model body line by line,
    materialising the enumerated tensor ``(|enum|, J, P, S, Sp)`` exactly as
    numpyro + funsor would, then summing / logsumexp-ing it.  No gradient.
``*_logp_grad``             closed-form per-site log-marginal and hand-derived
    reverse-mode gradient; this is the algorithm the CUDA kernels implement.

Reference lines restated (paths relative to /root/reference):
  biolith/models/occu.py:135-242      (occu body)
  biolith/models/occu_rn.py:123-222   (occu_rn body)
  biolith/models/occu_cop.py:146-255  (occu_cop body)
  biolith/models/nmixture.py:124-220  (nmixture body)
  biolith/models/occu_cs.py:120-223   (occu_cs body)
  biolith/regression/linear.py:16-66  (LinearRegression)
  biolith/utils/modeling.py:8-39      (mask_missing_obs / flatten / reshape)
  biolith/utils/distributions.py:6-40 (RightTruncatedPoisson)
  biolith/utils/data.py:113-140       (period-dim insertion, fp32 cast)

numpyro semantics restated (numpyro/distributions/{discrete,util,distribution}.py,
numpyro>=0.18; recalled, not vendored):
  clamp_probs(p)            = clip(p, finfo.tiny, 1 - finfo.eps)
  BernoulliProbs.log_prob   = xlogy(v, p~) + xlog1py(1 - v, -p~)
  Poisson.log_prob          = xlogy(v, rate) - gammaln(v + 1) - rate
  BinomialProbs.log_prob    = gammaln(n+1) - gammaln(v+1) - gammaln(n-v+1) + xlogy(v, p) + xlog1py(n-v, -p)
  CategoricalLogits         = logits - logsumexp(logits)   (normalised)
  MaskedDistribution        = where(m, base.log_prob(where(m, v, feasible)), 0)
  Normal.log_prob           = -((v - loc)^2) / (2 scale^2) - log(scale) - log(sqrt(2 pi))
  TruncatedDistribution(Normal(loc, s), low=a).log_prob(v) = Normal.log_prob(v) - log(1 - Phi((a - loc)/s));
                              support greater_than(a) -> biject_to = Exp then Affine(a, 1): v = a + exp(x)
                              (dynamic support: the transform is rebuilt from the traced `low` every call)
  Gamma(c, r).log_prob      = c log r + (c-1) log v - r v - gammaln(c);  positive support -> ExpTransform
  Beta / Exponential priors are sampled in unconstrained space through
  SigmoidTransform / ExpTransform with their log-Jacobians (potential_fn).

Parameter vector layout used everywhere in this repo (one chain):
  theta = [ beta (Kb = Ks+1) | alpha (Ka = Ko+1) | extras ]
  extras (unconstrained), in this order when enabled:
     occu     : logit(prob_fp_constant), logit(prob_fp_unoccupied)
     occu_rn  : logit(prob_fp_constant)
     occu_cop : log(rate_fp_constant), log(rate_fp_unoccupied)
     occu_cs  : mu0, log(mu1 - mu0), log(sigma0), log(sigma1)   (always present)
Only n_species == 1 is restated in closed form (the enumerated form handles Sp >= 1).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy.special import expit, gammaln, log_ndtr, xlog1py, xlogy

LOG_2PI = float(np.log(2.0 * np.pi))


def logsumexp(a, axis=None, keepdims=False):
    """Max-shifted log-sum-exp (jax.scipy.special.logsumexp semantics for finite maxima)."""
    a = np.asarray(a, np.float64)
    amax = np.max(a, axis=axis, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, 0.0)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(a - amax), axis=axis, keepdims=True)) + amax
    return out if keepdims else np.squeeze(out, axis=axis)


# --------------------------------------------------------------------------- data
@dataclass
class Prepared:
    """Output of the reference's NaN-mask resolution (occu.py:136-142)."""

    y: np.ndarray  # (Sp, S, P, J) float64, NaN where masked by covariates or missing
    X: np.ndarray  # (S, Ks) NaN -> 0
    W: np.ndarray  # (S, P, J, Ko) NaN -> 0
    T: np.ndarray  # (S, P, J) session duration (ones if absent)
    mask: np.ndarray  # (Sp, S, P, J) bool == isfinite(y)  (modeling.py:15-17)


def ensure_period_dim(site_covs, obs_covs, obs, session_duration=None):
    """biolith/utils/data.py:113-127 (_ensure_season_dim)."""
    if obs_covs is not None:
        if obs_covs.ndim == 2:
            obs_covs = obs_covs[:, :, None]
        if obs_covs.ndim == 3:
            obs_covs = obs_covs[:, None, :, :]
    if obs is not None and obs.ndim == 2:
        obs = obs[:, None, :]
    if session_duration is not None and session_duration.ndim == 2:
        session_duration = session_duration[:, None, :]
    return site_covs, obs_covs, obs, session_duration


def prepare(site_covs, obs_covs, obs, session_duration=None, dtype=np.float32) -> Prepared:
    """NaN-mask resolution, occu.py:136-142 == occu_rn.py:124-130 == occu_cop.py:151-157.

    ``dtype`` is the dtype the reference would hold the arrays in (fp32 unless
    jax_enable_x64; data.py:135-140) -- it only matters for nan_to_num's +-inf
    replacement values.  Arithmetic afterwards is float64.
    """
    site_covs = np.asarray(site_covs, dtype=dtype)
    obs_covs = np.asarray(obs_covs, dtype=dtype)
    obs = np.asarray(obs, dtype=dtype)
    assert obs.ndim == 4 and site_covs.ndim == 2 and obs_covs.ndim == 4
    S, P, J, _ = obs_covs.shape
    assert obs.shape[1:] == (S, P, J) and site_covs.shape[0] == S
    obs_mask = np.isnan(obs_covs).any(axis=-1) | np.isnan(site_covs).any(axis=-1)[:, None, None]
    obs = np.where(obs_mask[None, ...], np.nan, obs)
    obs_covs = np.nan_to_num(obs_covs)
    site_covs = np.nan_to_num(site_covs)
    if session_duration is None:
        T = np.ones((S, P, J), dtype=np.float64)  # occu_cop.py:146-148
    else:
        T = np.asarray(session_duration, dtype=np.float64)
    return Prepared(
        y=obs.astype(np.float64),
        X=site_covs.astype(np.float64),
        W=obs_covs.astype(np.float64),
        T=T,
        mask=np.isfinite(obs),
    )


def n_extras(model: str, fp_constant=False, fp_unoccupied=False) -> int:
    if model == "occu_cs":
        return 4
    if model == "occu_rn":
        assert not fp_unoccupied
    return int(bool(fp_constant)) + int(bool(fp_unoccupied))


def split_theta(theta, Ks, Ko, fp_constant, fp_unoccupied):
    theta = np.asarray(theta, dtype=np.float64)
    Kb, Ka = Ks + 1, Ko + 1
    beta = theta[:Kb]
    alpha = theta[Kb : Kb + Ka]
    i = Kb + Ka
    xc = xu = None
    if fp_constant:
        xc = theta[i]
        i += 1
    if fp_unoccupied:
        xu = theta[i]
        i += 1
    assert i == theta.shape[0], "theta has wrong length"
    return beta, alpha, xc, xu


# ---------------------------------------------------------------- numpyro pieces
def _finfo(dtype):
    return np.finfo(np.dtype(dtype))


def clamp_probs(p, finfo):
    return np.clip(p, finfo.tiny, 1.0 - finfo.eps)


def bernoulli_log_prob(p, v, finfo):
    pt = clamp_probs(p, finfo)
    return xlogy(v, pt) + xlog1py(1.0 - v, -pt)


def poisson_log_prob(rate, v):
    with np.errstate(divide="ignore", invalid="ignore"):
        return xlogy(v, rate) - gammaln(v + 1.0) - rate


def normal_log_prob(x, loc=0.0, scale=1.0):
    return -0.5 * ((x - loc) / scale) ** 2 - np.log(scale) - 0.5 * LOG_2PI


def _linear(coef, covs_flat):
    """LinearRegression.__call__, linear.py:48-66.  coef (Sp, k+1), covs (n, k) -> (n, Sp)."""
    intercept = coef[..., 0]
    slopes = coef[..., 1:]
    linear = np.tensordot(slopes, covs_flat, axes=([-1], [1]))  # (Sp, n)
    return linear.T + intercept.reshape((1,) + intercept.shape)


def _flatten(covs):
    """modeling.py:22-28: (n_covs, *obs_shape) -> (n_obs, n_covs), obs_shape."""
    # (explicit size instead of -1 so that an intercept-only predictor, n_covs = 0, stays well defined)
    return covs.reshape(covs.shape[0], int(np.prod(covs.shape[1:]))).T, covs.shape[1:]


def _extras_prior_sigmoid(x, a=2.0, b=5.0):
    """Beta(a,b) on c = sigmoid(x), plus log|dc/dx| (SigmoidTransform)."""
    c = expit(x)
    logc = -np.logaddexp(0.0, -x)
    log1mc = -np.logaddexp(0.0, x)
    lp = (a - 1.0) * logc + (b - 1.0) * log1mc + gammaln(a + b) - gammaln(a) - gammaln(b)
    return c, lp + logc + log1mc, a * (1.0 - c) - b * c  # value, log-density, d/dx


def _extras_prior_exp(x, rate=1.0):
    """Exponential(rate) on c = exp(x), plus log|dc/dx| = x (ExpTransform)."""
    c = np.exp(x)
    return c, np.log(rate) - rate * c + x, -rate * c + 1.0


# ===================================================================== enumerated
def occu_log_joint_enumerated(
    theta, site_covs, obs_covs, obs, *, false_positives_constant=False,
    false_positives_unoccupied=False, dtype=np.float32, prior=True,
):
    """occu.py:135-242 op by op; theta = [beta(Sp*Kb) | alpha(Sp*Ka) | extras]."""
    finfo = _finfo(dtype)
    pr = prepare(site_covs, obs_covs, obs, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    theta = np.asarray(theta, np.float64)
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    beta = theta[:nb].reshape(Sp, Ks + 1)
    alpha = theta[nb : nb + na].reshape(Sp, Ko + 1)
    rest = list(theta[nb + na :])
    lp = 0.0
    c = u = 0.0
    if false_positives_constant:  # occu.py:146-150, Beta(2,5) in logit space
        c, l, _ = _extras_prior_sigmoid(rest.pop(0))
        lp += l if prior else 0.0
    if false_positives_unoccupied:  # occu.py:153-157
        u, l, _ = _extras_prior_sigmoid(rest.pop(0))
        lp += l if prior else 0.0
    assert not rest
    if prior:  # linear.py:28, prior.expand([k+1]).to_event(1), default Normal(0,1)
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    # occu.py:176-180
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    y = pr.y.transpose(3, 2, 1, 0)  # (J,P,S,Sp)
    m = np.isfinite(y)
    occ_linear = _linear(beta, site_flat).reshape(site_shape + (Sp,))  # (S,Sp)
    psi = expit(occ_linear)  # occu.py:207 -> broadcast (P,S,Sp)
    psi = np.broadcast_to(psi, (P, S, Sp))
    z = np.array([0.0, 1.0]).reshape(2, 1, 1, 1, 1)  # enum dim = -5
    log_pz = bernoulli_log_prob(psi, z, finfo)  # (2,1,P,S,Sp)
    p = expit(_linear(alpha, obs_flat).reshape(obs_shape + (Sp,)))  # (J,P,S,Sp)
    p_fp = 1 - (1 - z * p) * (1 - c) * (1 - (1 - z) * u)  # occu.py:229-235
    v = np.where(m, y, 0.0)
    ll = np.where(m, bernoulli_log_prob(p_fp, v, finfo), 0.0)  # (2,J,P,S,Sp)
    site_ll = ll.sum(axis=1, keepdims=True) + log_pz  # (2,1,P,S,Sp)
    return float(lp + logsumexp(site_ll, axis=0).sum())


def occu_rn_log_joint_enumerated(
    theta, site_covs, obs_covs, obs, *, max_abundance=100, false_positives_constant=False,
    dtype=np.float32, prior=True,
):
    """occu_rn.py:123-222 op by op."""
    finfo = _finfo(dtype)
    pr = prepare(site_covs, obs_covs, obs, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    theta = np.asarray(theta, np.float64)
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    beta = theta[:nb].reshape(Sp, Ks + 1)
    alpha = theta[nb : nb + na].reshape(Sp, Ko + 1)
    rest = list(theta[nb + na :])
    lp = 0.0
    c = 0.0
    if false_positives_constant:
        c, l, _ = _extras_prior_sigmoid(rest.pop(0))
        lp += l if prior else 0.0
    assert not rest
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    y = pr.y.transpose(3, 2, 1, 0)
    m = np.isfinite(y)
    abu_linear = _linear(beta, site_flat).reshape(site_shape + (Sp,))
    abundance = np.broadcast_to(np.exp(abu_linear), (P, S, Sp))  # occu_rn.py:188
    # distributions.py:31-40: Categorical(logits=Poisson(rate[...,None]).log_prob(0..K))
    support = np.arange(max_abundance + 1, dtype=np.float64)
    logits = poisson_log_prob(abundance[..., None], support)  # (P,S,Sp,K+1)
    log_pN = logits - logsumexp(logits, axis=-1, keepdims=True)
    log_pN = np.moveaxis(log_pN, -1, 0)[:, None]  # (K+1,1,P,S,Sp)
    N = support.reshape(-1, 1, 1, 1, 1)
    r = expit(_linear(alpha, obs_flat).reshape(obs_shape + (Sp,)))  # (J,P,S,Sp)
    p_it = 1.0 - (1.0 - r) ** N  # occu_rn.py:213 -> (K+1,J,P,S,Sp)
    v = np.where(m, y, 0.0)
    ll = np.where(m, bernoulli_log_prob(1 - (1 - p_it) * (1 - c), v, finfo), 0.0)
    site_ll = ll.sum(axis=1, keepdims=True) + log_pN
    return float(lp + logsumexp(site_ll, axis=0).sum())


def occu_cop_log_joint_enumerated(
    theta, site_covs, obs_covs, obs, session_duration=None, *, false_positives_constant=False,
    false_positives_unoccupied=False, dtype=np.float32, prior=True,
):
    """occu_cop.py:146-255 op by op."""
    finfo = _finfo(dtype)
    pr = prepare(site_covs, obs_covs, obs, session_duration, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    theta = np.asarray(theta, np.float64)
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    beta = theta[:nb].reshape(Sp, Ks + 1)
    alpha = theta[nb : nb + na].reshape(Sp, Ko + 1)
    rest = list(theta[nb + na :])
    lp = 0.0
    c = u = 0.0
    if false_positives_constant:  # occu_cop.py:160-164, Exponential(1) in log space
        c, l, _ = _extras_prior_exp(rest.pop(0))
        lp += l if prior else 0.0
    if false_positives_unoccupied:  # occu_cop.py:167-171
        u, l, _ = _extras_prior_exp(rest.pop(0))
        lp += l if prior else 0.0
    assert not rest
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    T = pr.T.transpose(2, 1, 0)[..., None]  # occu_cop.py:192 -> (J,P,S,1)
    y = pr.y.transpose(3, 2, 1, 0)
    m = np.isfinite(y)
    psi = np.broadcast_to(expit(_linear(beta, site_flat).reshape(site_shape + (Sp,))), (P, S, Sp))
    z = np.array([0.0, 1.0]).reshape(2, 1, 1, 1, 1)
    log_pz = bernoulli_log_prob(psi, z, finfo)
    rate = np.exp(_linear(alpha, obs_flat).reshape(obs_shape + (Sp,)))
    l_det = z * rate + (1 - z) * u + c  # occu_cop.py:243-247
    v = np.where(m, y, 0.0)
    ll = np.where(m, poisson_log_prob(T * l_det, v), 0.0)
    site_ll = ll.sum(axis=1, keepdims=True) + log_pz
    return float(lp + logsumexp(site_ll, axis=0).sum())



def nmixture_log_joint_enumerated(theta, site_covs, obs_covs, obs, *, max_abundance=100, dtype=np.float32,
                                  prior=True):
    """nmixture.py:124-220 op by op (Royle 2004 N-mixture; truncated, un-renormalised Poisson prior)."""
    pr = prepare(site_covs, obs_covs, obs, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    theta = np.asarray(theta, np.float64)
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    beta = theta[:nb].reshape(Sp, Ks + 1)
    alpha = theta[nb : nb + na].reshape(Sp, Ko + 1)
    assert theta.size == nb + na
    lp = 0.0
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    y = pr.y.transpose(3, 2, 1, 0)  # (J,P,S,Sp)
    m = np.isfinite(y)
    # nmixture.py:167-171
    obs_max = np.max(np.where(np.isnan(y), -np.inf, y), axis=0)
    obs_max = np.where(np.isfinite(obs_max), obs_max, 0)
    min_counts = obs_max.astype(int)  # (P,S,Sp)
    abundance = np.broadcast_to(np.exp(_linear(beta, site_flat).reshape(site_shape + (Sp,))), (P, S, Sp))
    support = np.arange(max_abundance + 1)
    logits = poisson_log_prob(abundance[..., None], support.astype(np.float64))  # (P,S,Sp,K+1)
    logits = np.where(support < min_counts[..., None], -np.inf, logits)
    factor = logsumexp(logits, axis=-1)  # numpyro.factor("N_i_trunc_norm", ...), nmixture.py:187
    with np.errstate(invalid="ignore"):
        log_pN = logits - factor[..., None]  # Categorical(logits) normalises
    log_pN = np.moveaxis(log_pN, -1, 0)[:, None]  # (K+1,1,P,S,Sp)
    N = support.reshape(-1, 1, 1, 1, 1).astype(np.float64)
    p = expit(_linear(alpha, obs_flat).reshape(obs_shape + (Sp,)))  # (J,P,S,Sp)
    v = np.where(m, y, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        ll = (gammaln(N + 1) - gammaln(v + 1) - gammaln(N - v + 1) + xlogy(v, p) + xlog1py(N - v, -p))
        ll = np.where(N - v < 0, -np.inf, ll)  # gammaln(non-positive integer) = +inf
    ll = np.where(m, ll, 0.0)
    with np.errstate(invalid="ignore"):
        site_ll = ll.sum(axis=1, keepdims=True) + log_pN
    site_ll = np.where(np.isnan(site_ll), -np.inf, site_ll)
    return float(lp + logsumexp(site_ll, axis=0).sum() + factor.sum())


def _cs_extras(x, prior, prior_mu_scale, prior_sigma):
    """occu_cs.py:146-154 in unconstrained space: x = [mu0, log(mu1 - mu0), log sigma0, log sigma1].

    Returns (mu0, mu1, sigma0, sigma1), the log-prior incl. log-Jacobians, and d(log-prior)/dx holding the
    CONSTRAINED values fixed for the likelihood part (the caller chains the likelihood gradient)."""
    mu0, x1, xs0, xs1 = (float(v) for v in x)
    s = float(prior_mu_scale)
    a, b = (float(v) for v in prior_sigma)
    e1 = np.exp(x1)
    mu1 = mu0 + e1
    sg0, sg1 = np.exp(xs0), np.exp(xs1)
    lp = 0.0
    g = np.zeros(4)
    if prior:
        t = mu0 / s
        lp += normal_log_prob(mu0, 0.0, s)
        lp += normal_log_prob(mu1, 0.0, s) - log_ndtr(-t) + x1  # left-truncated at mu0, + log|d mu1/dx1|
        hazard = np.exp(-0.5 * t * t - 0.5 * LOG_2PI - log_ndtr(-t))  # phi(t) / (1 - Phi(t))
        g[0] = -mu0 / s**2 - mu1 / s**2 + hazard / s
        g[1] = -mu1 / s**2 * e1 + 1.0
        for i, (xs, sg) in enumerate(((xs0, sg0), (xs1, sg1))):
            lp += a * np.log(b) + (a - 1.0) * xs - b * sg - gammaln(a) + xs
            g[2 + i] = a - b * sg
    return (mu0, mu1, sg0, sg1), lp, g


def occu_cs_log_joint_enumerated(theta, site_covs, obs_covs, obs, *, dtype=np.float32, prior=True,
                                 prior_mu_scale=10.0, prior_sigma=(5.0, 1.0)):
    """occu_cs.py:120-223 op by op (continuous-score occupancy, Rhinehart et al. 2022): two enumerated
    latents, z per (period, site) and f per visit; theta = [beta(Sp*Kb) | alpha(Sp*Ka) | mu0, x1, xs0, xs1]."""
    finfo = _finfo(dtype)
    pr = prepare(site_covs, obs_covs, obs, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    theta = np.asarray(theta, np.float64)
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    beta = theta[:nb].reshape(Sp, Ks + 1)
    alpha = theta[nb : nb + na].reshape(Sp, Ko + 1)
    assert theta.size == nb + na + 4
    (mu0, mu1, sg0, sg1), lp, _ = _cs_extras(theta[nb + na :], prior, prior_mu_scale, prior_sigma)
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    y = pr.y.transpose(3, 2, 1, 0)  # (J,P,S,Sp)
    m = np.isfinite(y)
    psi = np.broadcast_to(expit(_linear(beta, site_flat).reshape(site_shape + (Sp,))), (P, S, Sp))
    z = np.array([0.0, 1.0]).reshape(1, 2, 1, 1, 1, 1)  # enum dim of z
    f = np.array([0.0, 1.0]).reshape(2, 1, 1, 1, 1, 1)  # enum dim of f (allocated after z)
    log_pz = bernoulli_log_prob(psi, z, finfo)  # (1,2,1,P,S,Sp)
    p = expit(_linear(alpha, obs_flat).reshape(obs_shape + (Sp,)))  # (J,P,S,Sp)
    log_pf = bernoulli_log_prob(z * p, f, finfo)  # (2,2,J,P,S,Sp), occu_cs.py:202-213
    v = np.where(m, y, 0.0)
    ll = normal_log_prob(v, (1 - f) * mu0 + f * mu1, (1 - f) * sg0 + f * sg1)  # occu_cs.py:215-223
    ll = np.where(m, ll, 0.0)
    visit = logsumexp(log_pf + ll, axis=0)  # sum out f inside the replicate plate -> (2,J,P,S,Sp)
    site_ll = visit.sum(axis=1, keepdims=True) + log_pz[0]  # (2,1,P,S,Sp)
    return float(lp + logsumexp(site_ll, axis=0).sum())


# ==================================================================== closed form
def _softplus(x):
    return np.logaddexp(0.0, x)


def _clamped_log_sigmoid_pair(x, finfo):
    """log(p~), log1p(-p~), and the in-range indicator for p = sigmoid(x), p~ = clamp_probs(p).

    Clamp decisions are taken in log space (exact): p < 1-eps <=> log(1-p) > log(eps) and
    p > tiny <=> log(p) > log(tiny), so fp64 rounding of p near 1 cannot flip them.
    """
    p = expit(x)
    lp = -_softplus(-x)
    l1mp = -_softplus(x)
    lo = lp <= np.log(finfo.tiny)
    hi = l1mp <= np.log(finfo.eps)
    inr = ~(lo | hi)
    logp = np.where(inr, lp, np.where(lo, np.log(finfo.tiny), np.log1p(-finfo.eps)))
    log1mp = np.where(inr, l1mp, np.where(lo, np.log1p(-finfo.tiny), np.log(finfo.eps)))
    return p, logp, log1mp, inr


def _bern_terms_from_log1mP(lq, y, m, finfo, dlq_dnu, dlq_dc):
    """Clamped Bernoulli log-prob of y given P = 1 - exp(lq), and its derivatives through lq.

    lq = log(1 - P) is exact (log space); P~ = clip(P, tiny, 1 - eps):
      y = 0: log1p(-P~) = lq            if tiny < P < 1-eps  (<=> lq > log eps)
      y = 1: log(P~)    = log(-expm1(lq))
    d/dx = dt/dlq * dlq/dx with dt/dlq = 1 (y=0) or -(1-P)/P (y=1), zero where clipped.
    """
    P = -np.expm1(lq)
    log_eps = np.log(finfo.eps)
    inr = (lq > log_eps) & (P > finfo.tiny)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_y0 = np.where(inr, lq, np.where(P <= finfo.tiny, np.log1p(-finfo.tiny), log_eps))
        t_y1 = np.where(inr, np.log(np.where(inr, P, 1.0)),
                        np.where(P <= finfo.tiny, np.log(finfo.tiny), np.log1p(-finfo.eps)))
        dt_dlq = np.where(inr, np.where(y > 0.5, -np.exp(lq) / np.where(inr, P, 1.0), 1.0), 0.0)
    t = np.where(m, np.where(y > 0.5, t_y1, t_y0), 0.0)
    dt_dlq = np.where(m, dt_dlq, 0.0)
    return t, dt_dlq * dlq_dnu, dt_dlq * dlq_dc


def _units(pr: Prepared):
    """Flatten (site, period) into independent units; species 0 only."""
    Sp, S, P, J = pr.y.shape
    assert Sp == 1, "closed form restated for n_species == 1"
    U = S * P
    X = np.repeat(pr.X, P, axis=0)  # unit u = s*P + p shares site s covariates
    W = pr.W.reshape(U, J, -1)
    T = pr.T.reshape(U, J)
    m = pr.mask[0].reshape(U, J)
    y = np.where(m, pr.y[0].reshape(U, J), 0.0)
    return X, W, T, y, m


def occu_logp_grad(
    theta, pr: Prepared, *, fp_constant=False, fp_unoccupied=False, dtype=np.float32,
    prior=True, return_site_terms=False,
):
    """Closed-form occu log-density + gradient.  theta: (D,)."""
    finfo = _finfo(dtype)
    X, W, T, y, m = _units(pr)
    Ks, Ko = X.shape[1], W.shape[2]
    beta, alpha, xc, xu = split_theta(theta, Ks, Ko, fp_constant, fp_unoccupied)
    c = u = 0.0
    lp = 0.0
    gx = []
    if fp_constant:
        c, l, dc = _extras_prior_sigmoid(xc)
        lp += l if prior else 0.0
    if fp_unoccupied:
        u, l, du = _extras_prior_sigmoid(xu)
        lp += l if prior else 0.0
    eta = beta[0] + X @ beta[1:]
    nu = alpha[0] + W @ alpha[1:]  # (U,J)
    psi, logpsi, log1mpsi, in_psi = _clamped_log_sigmoid_pair(eta, finfo)
    if not (fp_constant or fp_unoccupied):
        p, logp, log1mp, in_p = _clamped_log_sigmoid_pair(nu, finfo)
        t1 = np.where(m, y * logp + (1 - y) * log1mp, 0.0)
        dt1_dnu = np.where(m & in_p, y - p, 0.0)  # y(1-p) - (1-y)p
        dt1_dc = 0.0
        P0 = 0.0
    else:
        # "ideal arithmetic" form: log(1-P1) = log(1-p) + log(1-c) is carried in log space so
        # that no 1-(1-x) cancellation enters (the reference's fp64 run has that noise; the
        # clamp decisions are made on the exact quantities)
        p = expit(nu)
        lq = -_softplus(nu) + np.log1p(-c)
        t1, dt1_dnu, dt1_dc = _bern_terms_from_log1mP(lq, y, m, finfo, dlq_dnu=-p, dlq_dc=-1.0 / (1.0 - c))
        P0 = -np.expm1(np.log1p(-c) + np.log1p(-u))
    n1 = (m * y).sum(axis=1)
    n0 = (m * (1 - y)).sum(axis=1)
    P0t = clamp_probs(P0, finfo)
    in0 = (P0 > finfo.tiny) and (P0 < 1.0 - finfo.eps)
    L0 = n1 * np.log(P0t) + n0 * np.log1p(-P0t)
    dL0_dP0 = (n1 / P0t - n0 / (1.0 - P0t)) if in0 else np.zeros_like(n1)
    L1 = t1.sum(axis=1)
    a = logpsi + L1
    b = log1mpsi + L0
    ell = np.logaddexp(a, b)
    r = expit(a - b)
    d_eta = np.where(in_psi, r - psi, 0.0)
    g_beta = np.concatenate([[d_eta.sum()], X.T @ d_eta])
    d_nu = r[:, None] * dt1_dnu
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("uj,ujk->k", d_nu, W)])
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    grads = [g_beta, g_alpha]
    if fp_constant:
        dl_dc = (r * np.sum(dt1_dc, axis=1)).sum() + ((1 - r) * dL0_dP0).sum() * (1.0 - u)
        grads.append([dl_dc * c * (1 - c) + (dc if prior else 0.0)])
    if fp_unoccupied:
        dl_du = ((1 - r) * dL0_dP0).sum() * (1.0 - c)
        grads.append([dl_du * u * (1 - u) + (du if prior else 0.0)])
    logp = float(lp + ell.sum())
    grad = np.concatenate([np.atleast_1d(g) for g in grads]).astype(np.float64)
    if return_site_terms:
        return logp, grad, dict(ell=ell, r=r, psi=psi, eta=eta, nu=nu)
    return logp, grad


def occu_rn_logp_grad(
    theta, pr: Prepared, *, max_abundance=100, fp_constant=False, dtype=np.float32,
    prior=True, return_site_terms=False,
):
    """Closed-form Royle-Nichols log-density + gradient."""
    finfo = _finfo(dtype)
    X, W, T, y, m = _units(pr)
    Ks, Ko = X.shape[1], W.shape[2]
    beta, alpha, xc, _ = split_theta(theta, Ks, Ko, fp_constant, False)
    c = 0.0
    lp = 0.0
    if fp_constant:
        c, l, dc = _extras_prior_sigmoid(xc)
        lp += l if prior else 0.0
    K = int(max_abundance)
    k = np.arange(K + 1, dtype=np.float64)
    eta = beta[0] + X @ beta[1:]
    lam = np.exp(eta)
    logits = xlogy(k[None, :], lam[:, None]) - gammaln(k + 1.0)[None, :] - lam[:, None]  # (U,K+1)
    log_pi = logits - logsumexp(logits, axis=1, keepdims=True)
    pi = np.exp(log_pi)
    nu = alpha[0] + W @ alpha[1:]
    r = expit(nu)
    uj = -_softplus(nu)  # log(1-r)
    # log(1 - P_kj) = k*log(1-r_j) + log(1-c), exact in log space (see _bern_terms_from_log1mP)
    lq = k[None, None, :] * uj[:, :, None] + np.log1p(-c)  # (U,J,K+1)
    t, dt_dnu, dt_dc = _bern_terms_from_log1mP(
        lq, y[:, :, None], m[:, :, None], finfo,
        dlq_dnu=-(k[None, None, :] * r[:, :, None]), dlq_dc=-1.0 / (1.0 - c))
    A = log_pi + t.sum(axis=1)  # (U,K+1)
    ell = logsumexp(A, axis=1)
    w = np.exp(A - ell[:, None])
    d_eta = (w * k).sum(axis=1) - (pi * k).sum(axis=1)
    d_nu = np.einsum("uk,ujk->uj", w, dt_dnu)
    g_beta = np.concatenate([[d_eta.sum()], X.T @ d_eta])
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("uj,ujk->k", d_nu, W)])
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    grads = [g_beta, g_alpha]
    if fp_constant:
        dl_dc = np.einsum("uk,ujk->", w, dt_dc)
        grads.append([dl_dc * c * (1 - c) + (dc if prior else 0.0)])
    logp = float(lp + ell.sum())
    grad = np.concatenate([np.atleast_1d(g) for g in grads]).astype(np.float64)
    if return_site_terms:
        return logp, grad, dict(ell=ell, w=w, eta=eta, nu=nu)
    return logp, grad


def occu_cop_logp_grad(
    theta, pr: Prepared, *, fp_constant=False, fp_unoccupied=False, dtype=np.float32,
    prior=True, return_site_terms=False,
):
    """Closed-form count-detection occupancy log-density + gradient."""
    finfo = _finfo(dtype)
    X, W, T, y, m = _units(pr)
    Ks, Ko = X.shape[1], W.shape[2]
    beta, alpha, xc, xu = split_theta(theta, Ks, Ko, fp_constant, fp_unoccupied)
    c = u = 0.0
    lp = 0.0
    if fp_constant:
        c, l, dc = _extras_prior_exp(xc)
        lp += l if prior else 0.0
    if fp_unoccupied:
        u, l, du = _extras_prior_exp(xu)
        lp += l if prior else 0.0
    eta = beta[0] + X @ beta[1:]
    psi, logpsi, log1mpsi, in_psi = _clamped_log_sigmoid_pair(eta, finfo)
    nu = alpha[0] + W @ alpha[1:]
    mu = np.exp(nu)
    rho1 = mu + c
    rho0 = u + c
    lg = gammaln(y + 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = np.where(m, xlogy(y, T * rho1) - lg - T * rho1, 0.0)
        t0 = np.where(m, xlogy(y, T * rho0) - lg - T * rho0, 0.0)
        d1 = np.where(m, np.where(y > 0, y / rho1, 0.0) - T, 0.0)  # d t1 / d rho1
        d0 = np.where(m, (np.where(y > 0, y / rho0, 0.0) if rho0 > 0 else 0.0) - T, 0.0)
    a = logpsi + t1.sum(axis=1)
    b = log1mpsi + t0.sum(axis=1)
    ell = np.logaddexp(a, b)
    with np.errstate(invalid="ignore"):
        r = np.where(np.isneginf(b), 1.0, expit(a - b))
    d_eta = np.where(in_psi, r - psi, 0.0)
    d_nu = r[:, None] * d1 * mu
    g_beta = np.concatenate([[d_eta.sum()], X.T @ d_eta])
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("uj,ujk->k", d_nu, W)])
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    grads = [g_beta, g_alpha]
    s1 = d1.sum(axis=1)
    s0 = np.where(r < 1.0, d0.sum(axis=1), 0.0)
    if fp_constant:
        dl_dc = (r * s1 + (1 - r) * s0).sum()
        grads.append([dl_dc * c + (dc if prior else 0.0)])
    if fp_unoccupied:
        dl_du = ((1 - r) * s0).sum()
        grads.append([dl_du * u + (du if prior else 0.0)])
    logp = float(lp + ell.sum())
    grad = np.concatenate([np.atleast_1d(g) for g in grads]).astype(np.float64)
    if return_site_terms:
        return logp, grad, dict(ell=ell, r=r, psi=psi, eta=eta, nu=nu)
    return logp, grad



def nmixture_logp_grad(theta, pr: Prepared, *, max_abundance=100, dtype=np.float32, prior=True,
                       return_site_terms=False):
    """Closed-form N-mixture log-density + gradient.  Per unit, with u_j = log(1-p_j):
       A_k = k (eta + sum_j m_j u_j) - lambda - lgamma(k+1) + sum_j m_j log C(k, y_j),  k >= max_j y_j
       l   = sum_j m_j y_j nu_j + logsumexp_k A_k;   dl/deta = E_w[k] - lambda;  dl/dnu_j = m_j (y_j - p_j E_w[k])."""
    X, W, T, y, m = _units(pr)
    Ks, Ko = X.shape[1], W.shape[2]
    beta, alpha, _, _ = split_theta(theta, Ks, Ko, False, False)
    K = int(max_abundance)
    k = np.arange(K + 1, dtype=np.float64)
    eta = beta[0] + X @ beta[1:]
    lam = np.exp(eta)
    nu = alpha[0] + W @ alpha[1:]
    p = expit(nu)
    l1p = -_softplus(nu)
    U = (m * l1p).sum(axis=1)
    V = (m * y * nu).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        lgc = gammaln(k[None, None, :] + 1) - gammaln(y[:, :, None] + 1) - gammaln(k[None, None, :] - y[:, :, None] + 1)
        lgc = np.where(k[None, None, :] < y[:, :, None], -np.inf, lgc)
    cst = np.where(m[:, :, None], lgc, 0.0).sum(axis=1)  # (U,K+1) data-only
    A = k[None, :] * (eta + U)[:, None] - lam[:, None] - gammaln(k + 1.0)[None, :] + cst
    ls = logsumexp(A, axis=1)
    w = np.exp(A - ls[:, None])
    Ek = (w * k).sum(axis=1)
    ell = V + ls
    d_eta = Ek - lam
    d_nu = m * (y - p * Ek[:, None])
    g_beta = np.concatenate([[d_eta.sum()], X.T @ d_eta])
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("uj,ujk->k", d_nu, W)])
    lp = 0.0
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    logp = float(lp + ell.sum())
    grad = np.concatenate([g_beta, g_alpha]).astype(np.float64)
    if return_site_terms:
        return logp, grad, dict(ell=ell, w=w, eta=eta, nu=nu)
    return logp, grad


def occu_cs_logp_grad(theta, pr: Prepared, *, dtype=np.float32, prior=True, prior_mu_scale=10.0,
                      prior_sigma=(5.0, 1.0), return_site_terms=False):
    """Closed-form occu_cs log-density + gradient.  Per visit, with n0/n1 the two Normal log-densities of
    the score and q the clamped Bernoulli probability of f = 1 in the branch (z = 1: p~_j, z = 0: tiny):
       L_j = logaddexp(log(1-q) + n0_j, log q + n1_j),  w_j = P(f_j = 1 | s_j, z)
       l = logaddexp(log psi~ + sum_j m_j L_j(1), log(1-psi~) + sum_j m_j L_j(0)),  r = P(z = 1 | s)
       dL_j/dnu = w_j - p_j (z = 1, in range), dL_j/dmu_f = w_f e_f / sigma_f, dL_j/dlog sigma_f = w_f (e_f^2 - 1)."""
    finfo = _finfo(dtype)
    X, W, T, y, m = _units(pr)
    Ks, Ko = X.shape[1], W.shape[2]
    theta = np.asarray(theta, np.float64)
    assert theta.size == Ks + Ko + 6
    beta, alpha = theta[: Ks + 1], theta[Ks + 1 : Ks + Ko + 2]
    (mu0, mu1, sg0, sg1), lp, g_prior = _cs_extras(theta[Ks + Ko + 2 :], prior, prior_mu_scale, prior_sigma)
    eta = beta[0] + X @ beta[1:]
    nu = alpha[0] + W @ alpha[1:]
    psi, logpsi, log1mpsi, in_psi = _clamped_log_sigmoid_pair(eta, finfo)
    p, logq, log1mq, in_p = _clamped_log_sigmoid_pair(nu, finfo)
    e0, e1 = (y - mu0) / sg0, (y - mu1) / sg1
    n0 = -0.5 * e0 * e0 - np.log(sg0) - 0.5 * LOG_2PI
    n1 = -0.5 * e1 * e1 - np.log(sg1) - 0.5 * LOG_2PI
    mf = m.astype(np.float64)

    def branch(lq, l1q):
        a0, a1 = l1q + n0, lq + n1
        L = np.logaddexp(a0, a1)
        w1 = expit(a1 - a0)
        return (mf * L).sum(axis=1), mf * w1

    L1, w1 = branch(logq, log1mq)
    L0, v1 = branch(np.full_like(nu, np.log(finfo.tiny)), np.full_like(nu, np.log1p(-finfo.tiny)))
    a = logpsi + L1
    b = log1mpsi + L0
    ell = np.logaddexp(a, b)
    r = expit(a - b)
    d_eta = np.where(in_psi, r - psi, 0.0)
    d_nu = r[:, None] * np.where(in_p, w1 - mf * p, 0.0)
    g_beta = np.concatenate([[d_eta.sum()], X.T @ d_eta])
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("uj,ujk->k", d_nu, W)])
    wf1 = r[:, None] * w1 + (1 - r)[:, None] * v1            # P(f = 1 | s), mixed over z
    wf0 = mf - wf1
    g_mu0 = (wf0 * e0).sum() / sg0
    g_mu1 = (wf1 * e1).sum() / sg1
    g_xs0 = (wf0 * (e0 * e0 - 1.0)).sum()
    g_xs1 = (wf1 * (e1 * e1 - 1.0)).sum()
    g_ext = np.array([g_mu0 + g_mu1, g_mu1 * (mu1 - mu0), g_xs0, g_xs1]) + g_prior  # mu1 = mu0 + exp(x1)
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    logp = float(lp + ell.sum())
    grad = np.concatenate([g_beta, g_alpha, g_ext]).astype(np.float64)
    if return_site_terms:
        return logp, grad, dict(ell=ell, r=r, psi=psi, eta=eta, nu=nu)
    return logp, grad



# ============================================================ random effects (groundwork)
# SURVEY 8 row f4, third item: site / observation random effects of `occu` (occu.py:168-173, 191-196,
# 215-228).  NOT on the accelerated path yet (biolith_b200.fit raises for them); restated here so that the
# kernels of the next round have their checker.  theta layout (one chain, n_species = 1):
#   [ beta | alpha | log site_re_sd (if site) | log obs_re_sd (if obs) | a_s (S) | d_s (S) | o_spj (S*P*J) ]
# a = site_re_occ, d = site_re_det (both ~ Normal(0, site_re_sd), occu.py:191-193), o = obs_re ~ Normal(0,
# obs_re_sd) for EVERY (s, p, j), masked or not (the plate is not under mask_missing_obs, occu.py:215-218);
# the two scales ~ HalfNormal(1) in log space (ExpTransform, log|J| = x).
def _half_normal_log(x, scale=1.0):
    """HalfNormal(scale) on sd = exp(x) plus log|d sd/dx|; returns (sd, log-density, d/dx)."""
    sd = np.exp(x)
    lp = np.log(2.0) + normal_log_prob(sd, 0.0, scale) + x
    return sd, lp, -(sd / scale) ** 2 + 1.0


def occu_re_dims(S, P, J, Ks, Ko, site_re, obs_re):
    n_sd = int(bool(site_re)) + int(bool(obs_re))
    return Ks + Ko + 2 + n_sd + (2 * S if site_re else 0) + (S * P * J if obs_re else 0)


def _split_re(theta, S, P, J, Ks, Ko, site_re, obs_re):
    theta = np.asarray(theta, np.float64)
    i = Ks + Ko + 2
    beta, alpha = theta[: Ks + 1], theta[Ks + 1 : i]
    xs = xo = None
    if site_re:
        xs = theta[i]; i += 1
    if obs_re:
        xo = theta[i]; i += 1
    a = d = np.zeros(S)
    o = np.zeros((S, P, J))
    if site_re:
        a = theta[i : i + S]; i += S
        d = theta[i : i + S]; i += S
    if obs_re:
        o = theta[i : i + S * P * J].reshape(S, P, J); i += S * P * J
    assert i == theta.size, "theta has wrong length"
    return beta, alpha, xs, xo, a, d, o


def occu_re_log_joint_enumerated(theta, site_covs, obs_covs, obs, *, site_random_effects=False,
                                 obs_random_effects=False, dtype=np.float32, prior=True):
    """occu.py:135-242 with the random-effect branches, op by op (no false positives, n_species = 1)."""
    finfo = _finfo(dtype)
    pr = prepare(site_covs, obs_covs, obs, dtype=dtype)
    Sp, S, P, J = pr.y.shape
    assert Sp == 1
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    beta, alpha, xs, xo, a, d, o = _split_re(theta, S, P, J, Ks, Ko, site_random_effects, obs_random_effects)
    lp = 0.0
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
    if site_random_effects:
        sd, l, _ = _half_normal_log(xs)
        lp += l if prior else 0.0
        lp += normal_log_prob(a, 0.0, sd).sum() + normal_log_prob(d, 0.0, sd).sum()  # sample sites, always scored
    if obs_random_effects:
        sdo, l, _ = _half_normal_log(xo)
        lp += l if prior else 0.0
        lp += normal_log_prob(o, 0.0, sdo).sum()
    site_flat, site_shape = _flatten(pr.X.transpose(1, 0))
    obs_flat, obs_shape = _flatten(pr.W.transpose(3, 2, 1, 0))
    y = pr.y.transpose(3, 2, 1, 0)  # (J,P,S,1)
    m = np.isfinite(y)
    occ_linear = _linear(beta[None], site_flat).reshape(site_shape + (1,)) + a[:, None]  # (S,1)
    psi = np.broadcast_to(expit(occ_linear), (P, S, 1))
    z = np.array([0.0, 1.0]).reshape(2, 1, 1, 1, 1)
    log_pz = bernoulli_log_prob(psi, z, finfo)
    det_linear = (_linear(alpha[None], obs_flat).reshape(obs_shape + (1,)) + d[None, None, :, None]
                  + o.transpose(2, 1, 0)[..., None])  # (J,P,S,1)
    p = expit(det_linear)
    v = np.where(m, y, 0.0)
    ll = np.where(m, bernoulli_log_prob(z * p, v, finfo), 0.0)
    site_ll = ll.sum(axis=1, keepdims=True) + log_pz
    return float(lp + logsumexp(site_ll, axis=0).sum())


def occu_re_logp_grad(theta, pr: Prepared, *, site_random_effects=False, obs_random_effects=False,
                      dtype=np.float32, prior=True):
    """Closed form + gradient of occu with random effects: the per-unit terms of occu_logp_grad with
    eta_s + a_s and nu_spj + d_s + o_spj; the gradients w.r.t. a, d, o are ELEMENTWISE outputs."""
    finfo = _finfo(dtype)
    Sp, S, P, J = pr.y.shape
    assert Sp == 1
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    beta, alpha, xs, xo, a, d, o = _split_re(theta, S, P, J, Ks, Ko, site_random_effects, obs_random_effects)
    m = pr.mask[0]                                 # (S,P,J)
    y = np.where(m, pr.y[0], 0.0)
    eta = beta[0] + pr.X @ beta[1:] + a            # (S,)
    nu = alpha[0] + pr.W @ alpha[1:] + d[:, None, None] + o  # (S,P,J)
    psi, logpsi, log1mpsi, in_psi = _clamped_log_sigmoid_pair(eta, finfo)
    p, logp_, log1mp, in_p = _clamped_log_sigmoid_pair(nu, finfo)
    t1 = np.where(m, y * logp_ + (1 - y) * log1mp, 0.0)
    dt1 = np.where(m & in_p, y - p, 0.0)
    n1 = (m * y).sum(axis=2)
    n0 = (m * (1 - y)).sum(axis=2)
    L0 = n1 * np.log(finfo.tiny) + n0 * np.log1p(-finfo.tiny)
    av = logpsi[:, None] + t1.sum(axis=2)          # (S,P)
    bv = log1mpsi[:, None] + L0
    ell = np.logaddexp(av, bv)
    r = expit(av - bv)
    d_eta = np.where(in_psi[:, None], r - psi[:, None], 0.0)   # (S,P)
    d_nu = r[:, :, None] * dt1                                  # (S,P,J)
    g_beta = np.concatenate([[d_eta.sum()], pr.X.T @ d_eta.sum(axis=1)])
    g_alpha = np.concatenate([[d_nu.sum()], np.einsum("spj,spjk->k", d_nu, pr.W)])
    lp = ell.sum()
    if prior:
        lp += normal_log_prob(beta).sum() + normal_log_prob(alpha).sum()
        g_beta = g_beta - beta
        g_alpha = g_alpha - alpha
    grads = [g_beta, g_alpha]
    tail = []
    if site_random_effects:
        sd, l, dl = _half_normal_log(xs)
        lp += (l if prior else 0.0) + normal_log_prob(a, 0.0, sd).sum() + normal_log_prob(d, 0.0, sd).sum()
        # d/dx of sum_s [ -a^2/(2 sd^2) - log sd ] (x = log sd) = sum a^2/sd^2 - S, twice (a and d)
        grads.append([(a @ a + d @ d) / sd**2 - 2 * S + (dl if prior else 0.0)])
        tail += [d_eta.sum(axis=1) - a / sd**2, d_nu.sum(axis=(1, 2)) - d / sd**2]
    if obs_random_effects:
        sdo, l, dl = _half_normal_log(xo)
        lp += (l if prior else 0.0) + normal_log_prob(o, 0.0, sdo).sum()
        grads.append([(o * o).sum() / sdo**2 - o.size + (dl if prior else 0.0)])
        tail += [(d_nu - o / sdo**2).ravel()]
    grad = np.concatenate([np.atleast_1d(g) for g in grads + tail]).astype(np.float64)
    return float(lp), grad

# ------------------------------------------------------------------ conveniences
def logp_grad(model: str, theta, pr: Prepared, **kw):
    """Dispatch on model name; theta (D,) or (C, D) -> (logp[C], grad[C,D])."""
    fn = {"occu": occu_logp_grad, "occu_rn": occu_rn_logp_grad, "occu_cop": occu_cop_logp_grad,
          "nmixture": nmixture_logp_grad, "occu_cs": occu_cs_logp_grad}[model]
    theta = np.asarray(theta, np.float64)
    if theta.ndim == 1:
        return fn(theta, pr, **kw)
    out = [fn(t, pr, **kw) for t in theta]
    return np.array([o[0] for o in out]), np.stack([o[1] for o in out])


def log_joint_enumerated(model: str, theta, data: dict, **kw):
    fn = {
        "occu": occu_log_joint_enumerated,
        "occu_rn": occu_rn_log_joint_enumerated,
        "occu_cop": occu_cop_log_joint_enumerated,
        "nmixture": nmixture_log_joint_enumerated,
        "occu_cs": occu_cs_log_joint_enumerated,
    }[model]
    args = [data["site_covs"], data["obs_covs"], data["obs"]]
    if model == "occu_cop":
        args.append(data.get("session_duration"))
    return fn(theta, *args, **kw)


def finite_difference_grad(f, theta, h=1e-6):
    theta = np.asarray(theta, np.float64)
    g = np.zeros_like(theta)
    for i in range(theta.size):
        e = np.zeros_like(theta)
        e[i] = h * max(1.0, abs(theta[i]))
        g[i] = (f(theta + e) - f(theta - e)) / (2 * e[i])
    return g


def expected_mask(site_covs, obs_covs, obs) -> np.ndarray:
    """The bit-exact mask contract: (Sp,S,P,J) bool."""
    site_covs = np.asarray(site_covs)
    obs_covs = np.asarray(obs_covs)
    obs = np.asarray(obs)
    cov_nan = np.isnan(obs_covs).any(axis=-1) | np.isnan(site_covs).any(axis=-1)[:, None, None]
    return np.isfinite(obs) & ~cov_nan[None, ...]


def site_summary(model: str, thetas, pr: Prepared, **kw):
    """Per-unit posterior summaries over draws (checker for bl_site_summary): arrays of length S*P."""
    fn = {"occu": occu_logp_grad, "occu_rn": occu_rn_logp_grad, "occu_cop": occu_cop_logp_grad,
          "nmixture": nmixture_logp_grad, "occu_cs": occu_cs_logp_grad}[model]
    ells, a1, a2 = [], [], []
    for th in np.asarray(thetas, np.float64):
        _, _, t = fn(th, pr, prior=False, return_site_terms=True, **kw)
        ells.append(t["ell"])
        if model in ("occu_rn", "nmixture"):
            K = t["w"].shape[1] - 1
            a1.append(np.exp(t["eta"]))
            a2.append((t["w"] * np.arange(K + 1)).sum(axis=1))
        else:
            a1.append(t["psi"])
            a2.append(t["r"])
    ells = np.stack(ells)
    n = ells.shape[0]
    return dict(a1=np.mean(a1, axis=0), a2=np.mean(a2, axis=0), lppd=logsumexp(ells, axis=0) - np.log(n),
                p_waic=ells.var(axis=0, ddof=1) if n > 1 else np.zeros(ells.shape[1]))


# ------------------------------------------------------------------ per-observation outputs
def occu_deterministic_sites(thetas, pr: Prepared, obs_raw=None):
    """The deterministic sites the reference registers per draw (occu.py:207,221) in numpyro's own layout --
    psi (draws, S, Sp=1), prob_detection (draws, J, P, S, 1) -- and the reference's per-observation closed form
    `log_likelihood_manual` (evaluation/log_likelihood.py:55-98: psi*p form, clip(., 1e-10, 1 - 1e-10)),
    (draws, Sp, S, P, J), NaN where the observation is NaN."""
    thetas = np.atleast_2d(np.asarray(thetas, np.float64))
    Ks, Ko = pr.X.shape[1], pr.W.shape[3]
    eps = 1e-10
    psis, pdets, lls = [], [], []
    # the reference hands the RAW obs to log_likelihood_manual (only observation NaNs, not covariate NaNs)
    yT = np.asarray(obs_raw, np.float64).transpose(3, 2, 1, 0) if obs_raw is not None else None
    for th in thetas:
        beta, alpha = th[: Ks + 1], th[Ks + 1 : Ks + Ko + 2]
        psi = expit(beta[0] + pr.X @ beta[1:])[:, None]                        # (S, 1)
        pdet = expit(alpha[0] + pr.W @ alpha[1:]).transpose(2, 1, 0)[..., None]  # (J, P, S, 1)
        psis.append(psi)
        pdets.append(pdet)
        if yT is not None:
            q = pdet * psi[None, None]
            ll = np.log(np.clip(q, eps, 1 - eps)) * yT + np.log(np.clip(1 - q, eps, 1 - eps)) * (1 - yT)
            lls.append(ll.transpose(3, 2, 1, 0))
    return np.stack(psis), np.stack(pdets), (np.stack(lls) if lls else None)


def lppd_manual(log_lik_manual, data: dict) -> float:
    """evaluation/lppd.py:64-106: sum over valid observations of log-mean-exp over draws."""
    valid = (np.isfinite(data["obs"]) & np.isfinite(data["obs_covs"]).all(axis=-1)[None, ...]
             & np.isfinite(data["site_covs"]).all(axis=-1)[None, :, None, None])
    ll = np.asarray(log_lik_manual)[:, valid]
    return float(np.sum(logsumexp(ll, axis=0) - np.log(ll.shape[0])))
