"""jax.ffi binding of the accelerated path + drop-in NumPyro models.

UNTESTED IN THIS REPO'S CI: jax / numpyro / funsor are not installed (and not installable) in the
build image or on the GPU boxes, so this module is exercised only by its import guard
(tests/test_abi.py).  Everything below goes through the same C ABI (`bl_dataset_create`, `bl_eval`
via `bl_xla_eval`) that the ctypes tests cover on the GPU; the jax-facing part follows the
documented JAX >= 0.5 FFI API (`jax.ffi.register_ffi_target`, `jax.ffi.pycapsule`,
`jax.ffi.ffi_call`) and must be verified on a box that has jax ("verification").

Usage with the unmodified reference:

    from biolith.utils import fit                       # reference, unchanged
    from biolith_b200.jax_ffi import occu               # drop-in model, same signature / site names
    results = fit(occu, **data)                         # NUTS calls the B200 kernel per leapfrog

The drop-in models keep the reference's sample sites (`beta`, `alpha` through the unchanged
`LinearRegression.__init__`, biolith/regression/linear.py:16-28; `prob_fp_*` / `rate_fp_*`) so that
priors, `rename_samples` and every downstream consumer behave as before; only the plates +
enumeration + masked likelihood (occu.py:182-242) are replaced by
`numpyro.factor("loglik", occupancy_loglik(...))`.
"""

from __future__ import annotations

import ctypes as C
try:  # pragma: no cover - jax is absent in this image
    import jax
    import jax.numpy as jnp

    HAVE_JAX = True
except Exception:  # ModuleNotFoundError in this image
    jax = None
    jnp = None
    HAVE_JAX = False

from . import _lib
from .likelihood import OccupancyLikelihood

TARGET = "biolith_b200_eval"
_registered = False


def _require_jax():
    if not HAVE_JAX:
        raise ImportError(
            "biolith_b200.jax_ffi needs jax (and numpyro for the drop-in models); they are not installed. "
            "Use biolith_b200.fit / OccupancyLikelihood (ctypes path) instead.")


def register():
    """Register `bl_xla_eval` as an XLA custom-call target (legacy ABI, api_version=0)."""
    global _registered
    _require_jax()
    if not _registered:
        lib = _lib.load()
        jax.ffi.register_ffi_target(TARGET, jax.ffi.pycapsule(lib.bl_xla_eval), platform="CUDA", api_version=0)
        _registered = True


def make_loglik(likelihood: OccupancyLikelihood):
    """Returns `loglik(theta)` (theta: (..., D) on the GPU) with a custom VJP: one fused kernel launch
    yields value and gradient; the backward pass only scales the stored gradient by the cotangent."""
    _require_jax()
    register()
    D = likelihood.theta_dim
    dt = jnp.float32 if likelihood.dtype == "float32" else jnp.float64

    # chain batching: under vmap (numpyro chain_method="vectorized") ONE launch with a leading chain
    # axis, not C sequential calls -- the kernel is built around the chain batch
    @jax.custom_batching.custom_vmap
    def _call(theta2d):
        n = theta2d.shape[0]
        opaque = bytes(_lib.bl_xla_opaque(dataset=likelihood.handle, n_chains=n, reserved=0))
        out_types = (jax.ShapeDtypeStruct((n,), dt), jax.ShapeDtypeStruct((n, D), dt))
        return jax.ffi.ffi_call(TARGET, out_types, custom_call_api_version=2, legacy_backend_config=opaque,
                                vmap_method="sequential")(theta2d)

    @_call.def_vmap
    def _call_vmap(axis_size, in_batched, theta3d):
        n = theta3d.shape[1]
        lp, g = _call(theta3d.reshape(axis_size * n, D))
        return (lp.reshape(axis_size, n), g.reshape(axis_size, n, D)), (True, True)

    @jax.custom_vjp
    def loglik(theta):
        lp, _ = _call(theta.reshape(-1, D).astype(dt))
        return lp.reshape(theta.shape[:-1])

    def fwd(theta):
        lp, g = _call(theta.reshape(-1, D).astype(dt))
        return lp.reshape(theta.shape[:-1]), g.reshape(theta.shape)

    def bwd(g, ct):
        return (ct[..., None] * g,)

    loglik.defvjp(fwd, bwd)
    return loglik


def _drop_in(model_name):
    def model(site_covs, obs_covs, coords=None, ell=1.0, session_duration=None,
              false_positives_constant=False, false_positives_unoccupied=False, max_abundance=100, obs=None,
              n_species=1, prior_beta=None, prior_alpha=None, prior_mu=None, prior_sigma=None, **unsupported):
        _require_jax()
        import numpyro
        import numpyro.distributions as dist
        from biolith.regression import LinearRegression  # the reference's own regressor, unchanged

        if coords is not None or any(unsupported.get(k) for k in ("site_random_effects", "obs_random_effects")):
            raise _lib.BiolithB200Error(-2, model_name, "spatial / random effects are outside the accelerated path")
        if obs is None:
            raise _lib.BiolithB200Error(-2, model_name, "prior predictive (obs=None) is outside the accelerated path")
        lk = _handle_cache(model_name, site_covs, obs_covs, obs, session_duration, false_positives_constant,
                           false_positives_unoccupied, max_abundance)
        loglik = make_loglik(lk)
        extras = []
        if model_name == "occu_cs":
            # occu_cs.py:146-154, unchanged sample sites; the kernel takes the extras in numpyro's own
            # unconstrained coordinates, so jax differentiates through these three elementary maps
            pm = prior_mu if isinstance(prior_mu, tuple) else (prior_mu or dist.Normal(0, 10),) * 2
            ps = prior_sigma if isinstance(prior_sigma, tuple) else (prior_sigma or dist.Gamma(5, 1),) * 2
            mu0 = numpyro.sample("mu0", pm[0])
            mu1 = numpyro.sample("mu1", dist.TruncatedDistribution(pm[1], low=mu0))
            sigma0 = numpyro.sample("sigma0", ps[0])
            sigma1 = numpyro.sample("sigma1", ps[1])
            extras += [mu0, jnp.log(mu1 - mu0), jnp.log(sigma0), jnp.log(sigma1)]
        elif model_name == "occu_cop":
            if false_positives_constant:
                extras.append(jnp.log(numpyro.sample("rate_fp_constant", dist.Exponential())))
            if false_positives_unoccupied:
                extras.append(jnp.log(numpyro.sample("rate_fp_unoccupied", dist.Exponential())))
        else:
            if false_positives_constant:
                extras.append(jax.scipy.special.logit(numpyro.sample("prob_fp_constant", dist.Beta(2, 5))))
            if false_positives_unoccupied:
                extras.append(jax.scipy.special.logit(numpyro.sample("prob_fp_unoccupied", dist.Beta(2, 5))))
        with numpyro.plate("species", 1, dim=-1):
            reg_occ = LinearRegression("beta", site_covs.shape[1], prior=prior_beta or dist.Normal())
            reg_det = LinearRegression("alpha", obs_covs.shape[-1], prior=prior_alpha or dist.Normal())
        theta = jnp.concatenate([reg_occ.coef[0], reg_det.coef[0]] + [jnp.atleast_1d(e) for e in extras])
        numpyro.factor("loglik", loglik(theta))

    model.__name__ = model_name
    return model


_handles = {}


def _handle_cache(model_name, site_covs, obs_covs, obs, session_duration, fpc, fpu, max_abundance):
    """One packed dataset per (model, data identity): packing happens once per fit, like the
    reference's trace-time constant folding of the NaN mask (occu.py:136-142)."""
    import numpy as np

    key = (model_name, id(site_covs), id(obs_covs), id(obs), fpc, fpu, max_abundance)
    if key not in _handles:
        _handles[key] = OccupancyLikelihood(
            model_name, np.asarray(site_covs), np.asarray(obs_covs), np.asarray(obs),
            None if session_duration is None else np.asarray(session_duration),
            false_positives_constant=fpc, false_positives_unoccupied=fpu, max_abundance=max_abundance,
            prior=False)  # priors stay with numpyro's sample sites
    return _handles[key]


occu = _drop_in("occu")
occu_rn = _drop_in("occu_rn")
occu_cop = _drop_in("occu_cop")
nmixture = _drop_in("nmixture")
occu_cs = _drop_in("occu_cs")
