"""jax.ffi binding of the accelerated path + drop-in NumPyro models.

jax / numpyro / funsor are not installed (and not installable) in the build image or on the GPU boxes.
What IS exercised on the GPU (tests/test_gpu_xla_boundary.py): `bl_xla_eval` called exactly as XLA's legacy
custom-call ABI calls it (non-default stream, `void* buffers[3]`, opaque bytes, status + failure callback), and
this module's registration / custom_vjp / custom_vmap folding / drop-in models run against a minimal stand-in
for the handful of jax + numpyro entry points they use (tests/fakejax.py + oracle/refshim.py), with the result
compared to the executed reference bodies.  What remains unverified is only that the installed JAX accepts these
calls as documented for JAX >= 0.5 (`jax.ffi.register_ffi_target`, `jax.ffi.pycapsule`, `jax.ffi.ffi_call`);
see INTEGRATION.md, "verification".

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
from collections import OrderedDict

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
    loglik.batched_call = _call  # the custom_vmap'd primitive (tests fold a chain axis through it)
    return loglik


def _drop_in(model_name):
    def model(site_covs, obs_covs, coords=None, ell=1.0, session_duration=None, obs=None, **kwargs):
        _require_jax()
        import numpyro
        import numpyro.distributions as dist
        from biolith.regression import LinearRegression  # the reference's own regressor, unchanged

        from .models import model_options

        if obs is None:
            raise _lib.BiolithB200Error(-2, model_name, "prior predictive (obs=None) is outside the accelerated path")
        # same whitelist as biolith_b200.fit: unknown keys / options outside the path raise.  Priors are NOT
        # restricted here: every prior stays a numpyro sample site below, the kernel is likelihood-only
        from .models import KEYWORDS

        unknown = set(kwargs) - KEYWORDS[model_name]
        if unknown:
            raise _lib.BiolithB200Error(-1, model_name, f"unknown keyword(s): {sorted(unknown)}")
        fpc, fpu, max_abundance, _ = model_options(
            model_name, dict({k: v for k, v in kwargs.items() if not k.startswith("prior_")}, coords=coords))
        if obs.shape[0] != 1:
            raise _lib.BiolithB200Error(-2, model_name, "n_species > 1: one drop-in model per species")
        lk = dataset_handle(model_name, site_covs, obs_covs, obs, session_duration, fpc, fpu, max_abundance)
        loglik = make_loglik(lk)
        extras = []
        # the extras keep the reference's sample sites AND its priors (whatever distribution the caller passes:
        # the prior stays with numpyro); the kernel takes them in numpyro's own unconstrained coordinates, so jax
        # differentiates through these elementary maps
        if model_name == "occu_cs":  # occu_cs.py:146-154
            pm, ps = kwargs.get("prior_mu", dist.Normal(0, 10)), kwargs.get("prior_sigma", dist.Gamma(5, 1))
            pm = pm if isinstance(pm, tuple) else (pm, pm)
            ps = ps if isinstance(ps, tuple) else (ps, ps)
            mu0 = numpyro.sample("mu0", pm[0])
            mu1 = numpyro.sample("mu1", dist.TruncatedDistribution(pm[1], low=mu0))
            sigma0 = numpyro.sample("sigma0", ps[0])
            sigma1 = numpyro.sample("sigma1", ps[1])
            extras += [mu0, jnp.log(mu1 - mu0), jnp.log(sigma0), jnp.log(sigma1)]
        elif model_name == "occu_cop":  # occu_cop.py:160-171
            if fpc:
                extras.append(jnp.log(numpyro.sample(
                    "rate_fp_constant", kwargs.get("prior_rate_fp_constant", dist.Exponential()))))
            if fpu:
                extras.append(jnp.log(numpyro.sample(
                    "rate_fp_unoccupied", kwargs.get("prior_rate_fp_unoccupied", dist.Exponential()))))
        else:  # occu.py:146-157, occu_rn.py:133-137
            if fpc:
                extras.append(jax.scipy.special.logit(numpyro.sample(
                    "prob_fp_constant", kwargs.get("prior_prob_fp_constant", dist.Beta(2, 5)))))
            if fpu:
                extras.append(jax.scipy.special.logit(numpyro.sample(
                    "prob_fp_unoccupied", kwargs.get("prior_prob_fp_unoccupied", dist.Beta(2, 5)))))
        X = jnp.nan_to_num(jnp.asarray(site_covs))  # occu.py:141-142
        W = jnp.nan_to_num(jnp.asarray(obs_covs))
        with numpyro.plate("species", 1, dim=-1):
            reg_occ = LinearRegression("beta", site_covs.shape[1], prior=kwargs.get("prior_beta", dist.Normal()))
            reg_det = LinearRegression("alpha", obs_covs.shape[-1], prior=kwargs.get("prior_alpha", dist.Normal()))
            # deterministic sites of the reference with its shapes (occu.py:207,221; occu_rn.py:188,205;
            # occu_cop.py:222,236): (S, Sp) and (J, P, S, Sp).  Plain jnp -- evaluated once per kept draw by
            # numpyro's postprocess_fn, never per leapfrog (XLA dead-code-eliminates them from potential_fn)
            occ_linear = reg_occ(X)                                                     # (S, Sp)
            det_linear = reg_det(W.transpose((3, 2, 1, 0)).reshape(W.shape[3], -1).T)   # (J*P*S, Sp)
            det_linear = det_linear.reshape(W.shape[2], W.shape[1], W.shape[0], -1)
            if model_name in ("occu_rn", "nmixture"):
                numpyro.deterministic("abundance", jnp.exp(occ_linear))
            else:
                numpyro.deterministic("psi", jax.nn.sigmoid(occ_linear))
            if model_name == "occu_cop":
                numpyro.deterministic("rate_detection", jnp.exp(det_linear))
            elif model_name != "occu_cs":
                numpyro.deterministic("prob_detection", jax.nn.sigmoid(det_linear))
        theta = jnp.concatenate([reg_occ.coef[0], reg_det.coef[0]] + [jnp.atleast_1d(e) for e in extras])
        numpyro.factor("loglik", loglik(theta))

    model.__name__ = model_name
    return model


# ---- packed datasets: one per distinct (model, options, DATA CONTENT) --------------------------------------
# The reference's fit() builds fresh arrays on every call, so object identity says nothing: the key is a digest of
# the bytes.  A small LRU bounds the HBM held by finished fits; close_datasets() frees everything explicitly.
_handles = OrderedDict()
MAX_CACHED_DATASETS = 4


def _digest(*arrays) -> str:
    import hashlib

    import numpy as np

    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"<none>")
            continue
        a = np.ascontiguousarray(np.asarray(a))
        h.update(str((a.shape, a.dtype.str)).encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def dataset_handle(model_name, site_covs, obs_covs, obs, session_duration, fpc, fpu, max_abundance):
    """Packing happens once per distinct dataset, like the reference's trace-time constant folding of the NaN
    mask (occu.py:136-142); the model function itself is re-traced by numpyro several times per fit."""
    import numpy as np

    key = (model_name, bool(fpc), bool(fpu), int(max_abundance),
           _digest(site_covs, obs_covs, obs, session_duration if model_name == "occu_cop" else None))
    lk = _handles.get(key)
    if lk is not None:
        _handles.move_to_end(key)
        return lk
    lk = OccupancyLikelihood(
        model_name, np.asarray(site_covs), np.asarray(obs_covs), np.asarray(obs),
        None if session_duration is None else np.asarray(session_duration),
        false_positives_constant=fpc, false_positives_unoccupied=fpu, max_abundance=max_abundance,
        prior=False)  # priors stay with numpyro's sample sites
    _handles[key] = lk
    while len(_handles) > MAX_CACHED_DATASETS:
        _, old = _handles.popitem(last=False)
        old.close()
    return lk


def close_datasets():
    """Free every cached packed dataset (HBM) -- call after the last fit on a dataset."""
    while _handles:
        _, lk = _handles.popitem()
        lk.close()


occu = _drop_in("occu")
occu_rn = _drop_in("occu_rn")
occu_cop = _drop_in("occu_cop")
nmixture = _drop_in("nmixture")
occu_cs = _drop_in("occu_cs")
