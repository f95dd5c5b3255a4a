"""``fit`` with the signature of ``biolith.utils.fit`` (biolith/utils/fit.py:16-135), driving the
B200 kernels + device-resident NUTS instead of numpyro's MCMC.

Same positional/keyword surface (``model_fn, site_covs, obs_covs, obs, session_duration,
num_samples, num_warmup, random_seed, num_chains, kernel, init_strategy, timeout, **kwargs``), same
``FitResult(samples, mcmc)`` return, same sample names after ``rename_samples``
(biolith/utils/data.py:145-165: ``cov_state_<name>``, ``cov_det_<name>``).  Anything outside the
accelerated path (non-NUTS kernels, spatial / random effects, non-linear regressors, custom
non-Normal priors) raises ``BiolithB200Error`` -- there is no fallback.
"""

from __future__ import annotations

from collections import namedtuple
from typing import Callable, Optional

import numpy as np

from . import diagnostics as _diag
from ._lib import BiolithB200Error
from .likelihood import OccupancyLikelihood, _as_numpy
from .data import prepare_data, rename_samples
from .models import SUPPORTED, model_options
from .nuts import NutsSampler

FitResult = namedtuple("FitResult", ["samples", "mcmc"])

_MAX_DETERMINISTIC_ELEMS = 50_000_000


class MCMCResult:
    """Minimal stand-in for the ``numpyro.infer.MCMC`` object the reference returns."""

    def __init__(self, model, grouped, extra, num_samples, num_chains, info):
        self.model = model
        self._grouped = grouped  # name -> (chains, draws, ...)
        self._extra = extra
        self.num_samples, self.num_chains = num_samples, num_chains
        self.info = info

    def get_samples(self, group_by_chain=False):
        if group_by_chain:
            return dict(self._grouped)
        return {k: v.reshape((-1,) + v.shape[2:]) for k, v in self._grouped.items()}

    def get_extra_fields(self, group_by_chain=False):
        if group_by_chain:
            return dict(self._extra)
        return {k: v.reshape(-1) for k, v in self._extra.items()}

    def summary(self, prob=0.9):
        return _diag.summary(self._grouped, prob=prob)

    def print_summary(self, prob=0.9):
        s = self.summary(prob)
        print(f"{'':>24s} {'mean':>10s} {'std':>10s} {'median':>10s} {'n_eff':>10s} {'r_hat':>8s}")
        for name, d in s.items():
            mean, std, med, ne, rh = (np.atleast_1d(d[k]).ravel() for k in ("mean", "std", "median", "n_eff", "r_hat"))
            for i in range(mean.size):
                print(f"{name + ('[%d]' % i if mean.size > 1 else ''):>24s} {mean[i]:10.4f} {std[i]:10.4f} "
                      f"{med[i]:10.4f} {ne[i]:10.1f} {rh[i]:8.3f}")
        div = int(self._extra["diverging"].sum())
        print(f"Number of divergences: {div}")


def fit(
    model_fn: Callable,
    site_covs=None,
    obs_covs=None,
    obs=None,
    session_duration=None,
    num_samples: int = 1000,
    num_warmup: int = 1000,
    random_seed: int = 0,
    num_chains: int = 5,
    kernel: Optional[str] = None,
    init_strategy: Optional[Callable] = None,
    timeout: Optional[int] = None,
    *,
    dtype: str = "float32",
    device: int = 0,
    max_tree_depth: int = 10,
    target_accept_prob: float = 0.8,
    **kwargs,
) -> FitResult:
    name = getattr(model_fn, "__name__", str(model_fn))
    if name not in SUPPORTED:
        raise BiolithB200Error(-2, "fit", f"model {name!r} is outside the accelerated path {sorted(SUPPORTED)}")
    if kernel not in (None, "nuts"):
        raise BiolithB200Error(-2, "fit", f"kernel={kernel!r}: only NUTS is implemented on the device")
    if init_strategy is not None and init_strategy != "map":
        raise BiolithB200Error(-2, "fit", "init_strategy: only None (init_to_uniform(radius=2), the reference's default) "
                               "or \"map\" (start at the posterior mode, biolith_b200.optim) are supported")
    fpc, fpu, max_abundance, prior_kw = model_options(name, kwargs)
    site_covs, obs_covs, obs, session_duration, site_names, obs_names = prepare_data(
        site_covs, obs_covs, obs, session_duration)
    if site_covs is None or obs_covs is None or obs is None:
        raise BiolithB200Error(-1, "fit", "site_covs, obs_covs and obs are required (prior predictive is not accelerated)")
    from .likelihood import ensure_period_dim

    _, _, obs_np4, _ = ensure_period_dim(None, None, _as_numpy(obs), None)
    n_sp = obs_np4.shape[0]
    n_periods = obs_np4.shape[2]
    has_re = bool(prior_kw.get("site_random_effects") or prior_kw.get("obs_random_effects"))
    if n_sp > 1 and has_re:
        raise BiolithB200Error(-2, "fit", "random effects with n_species > 1 are outside the accelerated path")
    parts = []
    if n_sp > 1 and (fpc or fpu or name == "occu_cs"):
        # the false-positive / score parameters are sampled once, outside the species plate (occu.py:146-157,
        # occu_cs.py:146-154): they couple the species -> ONE composite handle, one joint NUTS run (csrc/multi.cu)
        parts.append(_fit_one(name, site_covs, obs_covs, obs_np4, session_duration, fpc, fpu, max_abundance,
                              dtype, device, num_chains, num_warmup, num_samples, random_seed, max_tree_depth,
                              target_accept_prob, timeout, prior_kw, n_periods, init_strategy))
        n_sp = 1  # the single part already carries every species
    for sp in range(n_sp if not parts else 0):
        # species are independent problems sharing the covariates (one handle each, occu.py:182-186)
        parts.append(_fit_one(name, site_covs, obs_covs, obs_np4[sp:sp + 1], session_duration, fpc, fpu, max_abundance,
                              dtype, device, num_chains, num_warmup, num_samples, random_seed + 7919 * sp,
                              max_tree_depth, target_accept_prob, timeout, prior_kw, n_periods,
                              init_strategy))
    grouped = {}
    for k in parts[0][0]:
        ax = 2 if k in ("beta", "alpha") else -1
        if k in ("beta", "alpha", "psi", "abundance"):
            grouped[k] = np.concatenate([p_[0][k] for p_ in parts], axis=ax)
        else:
            grouped[k] = parts[0][0][k]
    extra = {k: (np.any([p_[1][k] for p_ in parts], axis=0) if k == "diverging"
                 else np.mean([p_[1][k] for p_ in parts], axis=0) if k == "accept_prob"
                 else np.sum([p_[1][k] for p_ in parts], axis=0))
             for k in parts[0][1]}
    info = parts[0][2] if n_sp == 1 else {"per_species": [p_[2] for p_ in parts]}
    mcmc = MCMCResult(name, grouped, extra, num_samples, num_chains, info)
    samples = mcmc.get_samples()
    samples = rename_samples(samples, site_names, obs_names)
    return FitResult(samples, mcmc)


def _fit_one(name, site_covs, obs_covs, obs, session_duration, fpc, fpu, max_abundance, dtype, device, num_chains,
             num_warmup, num_samples, seed, max_tree_depth, target_accept_prob, timeout, prior_kw, n_periods,
             init_strategy=None):
    """One species: pack, sample on the device, return (grouped samples, extra fields, info)."""
    lk = OccupancyLikelihood(
        name, site_covs, obs_covs, obs, session_duration, false_positives_constant=fpc,
        false_positives_unoccupied=fpu, max_abundance=max_abundance, dtype=dtype, prior=True,
        device=device, max_chains=num_chains, **prior_kw)
    init_params = None
    if init_strategy == "map":
        from .optim import find_map, init_around

        theta_map, _, _ = find_map(lk, seed=seed)
        init_params = init_around(theta_map, num_chains, seed=seed)
    sampler = NutsSampler(lk, num_chains, num_warmup, num_samples, seed=seed, init_params=init_params,
                          max_tree_depth=max_tree_depth, target_accept_prob=target_accept_prob)
    complete = sampler.run(timeout=timeout)
    if not complete:
        sampler.close()
        lk.close()
        raise TimeoutError("sampling did not finish within the timeout")  # reference: utils/misc.py:11-21
    res = sampler.results()
    sampler.close()

    th = res["samples"].astype(np.float64)  # (C, N, D)
    Ks, Ko = lk.shape["n_site_covs"], lk.shape["n_obs_covs"]
    Sp = lk.n_species  # theta = [beta (Sp x Kb) | alpha (Sp x Ka) | extras]; Sp = 1 unless the species are coupled
    nb, na = Sp * (Ks + 1), Sp * (Ko + 1)
    grouped = {
        "beta": th[:, :, :nb].reshape(th.shape[0], th.shape[1], Sp, Ks + 1),  # (C, N, n_species, Kb): numpyro's layout
        "alpha": th[:, :, nb : nb + na].reshape(th.shape[0], th.shape[1], Sp, Ko + 1),
    }
    i = nb + na
    if name == "occu_cs":  # occu_cs.py:148-154; mu1 = mu0 + exp(x) (left-truncated at mu0)
        grouped["mu0"] = th[:, :, i]
        grouped["mu1"] = th[:, :, i] + np.exp(th[:, :, i + 1])
        grouped["sigma0"] = np.exp(th[:, :, i + 2])
        grouped["sigma1"] = np.exp(th[:, :, i + 3])
    elif name == "occu_cop":
        if fpc:
            grouped["rate_fp_constant"] = np.exp(th[:, :, i]); i += 1
        if fpu:
            grouped["rate_fp_unoccupied"] = np.exp(th[:, :, i]); i += 1
    else:
        if fpc:
            grouped["prob_fp_constant"] = 1 / (1 + np.exp(-th[:, :, i])); i += 1
        if fpu:
            grouped["prob_fp_unoccupied"] = 1 / (1 + np.exp(-th[:, :, i])); i += 1
    X = np.nan_to_num(np.asarray(_as_numpy(site_covs), dtype=np.float64))
    S = X.shape[0]
    a_re = 0.0
    if lk.site_random_effects or lk.obs_random_effects:
        # theta = [beta | alpha | log sd_site | log sd_obs | a (S) | d (S) | o (S*P*J)]; numpyro's site layouts:
        # site_re_* (S, Sp) inside plates site/species, obs_re (J, P, S, Sp) (occu.py:191-196, 215-218)
        P_, J_ = lk.shape["n_periods"], lk.shape["n_replicates"]
        if lk.site_random_effects:
            grouped["site_re_sd"] = np.exp(th[:, :, i]); i += 1
        if lk.obs_random_effects:
            grouped["obs_re_sd"] = np.exp(th[:, :, i]); i += 1
        if lk.site_random_effects:
            a_re = th[:, :, i : i + S]
            grouped["site_re_occ"] = a_re[..., None]; i += S
            grouped["site_re_det"] = th[:, :, i : i + S][..., None]; i += S
        if lk.obs_random_effects:
            o = th[:, :, i : i + S * P_ * J_].reshape(th.shape[0], th.shape[1], S, P_, J_)
            grouped["obs_re"] = o.transpose(0, 1, 4, 3, 2)[..., None]; i += S * P_ * J_
    # small problems: also materialise the deterministic site the reference's tests read
    if S * num_chains * num_samples <= _MAX_DETERMINISTIC_ELEMS:
        b = grouped["beta"]  # (C, N, Sp, Kb)
        eta = b[..., 0][:, :, None, :] + np.einsum("cnpk,sk->cnsp", b[..., 1:], X)  # (C, N, S, Sp)
        if not np.isscalar(a_re):
            eta = eta + a_re[..., None]
        det = np.exp(eta) if name in ("occu_rn", "nmixture") else 1 / (1 + np.exp(-eta))
        grouped["abundance" if name in ("occu_rn", "nmixture") else "psi"] = np.broadcast_to(
            det[:, :, None, :, :], det.shape[:2] + (n_periods, S, Sp))  # (C,N,P,S,Sp)
    extra = dict(diverging=res["diverging"], accept_prob=res["accept_prob"], num_steps=res["num_steps"],
                 potential_energy=res["potential_energy"])
    info = dict(step_size=res["step_size"], inverse_mass_matrix=res["inverse_mass_matrix"],
                leapfrogs=res["leapfrogs"], warmup_leapfrogs=res["warmup_leapfrogs"],
                global_steps=res["global_steps"], wall_s=res["wall_s"], kernel_variant=lk.kernel_variant)
    # per-site posterior summaries (psi / occupancy probability / pointwise lppd, p_waic), streamed on the
    # GPU over <= 512 thinned draws: available at any n_sites, unlike the per-draw deterministic sites
    if not (lk.site_random_effects or lk.obs_random_effects or Sp > 1):
        flat = th.reshape(-1, th.shape[-1])
        thin = flat[:: max(1, flat.shape[0] // 512)][:512]
        info["site_summary"] = lk.site_summary(thin)
    lk.close()
    return grouped, extra, info
