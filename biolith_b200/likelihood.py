"""Host-side handle of the accelerated path: one packed dataset on one GPU + logp/grad evaluation.

Mirrors the data contract of the reference models (biolith/models/occu.py:51-58):
``site_covs (S,Ks)``, ``obs_covs (S,P,J,Ko)``, ``obs (1,S,P,J)``, ``session_duration (S,P,J)``;
lower-rank inputs get the period dimension inserted exactly like biolith/utils/data.py:113-127.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import BiolithB200Error, bl_desc, bl_info, check

_DT = {"float32": (_lib.BL_F32, np.float32), "float64": (_lib.BL_F64, np.float64)}


def _as_numpy(a):
    if a is None:
        return None
    if hasattr(a, "to_numpy"):  # pandas
        a = a.to_numpy()
    return np.asarray(a)


def ensure_period_dim(site_covs, obs_covs, obs, session_duration=None):
    """Same rank promotion as the reference's ``_ensure_season_dim`` (utils/data.py:113-127)."""
    if obs_covs is not None:
        if obs_covs.ndim == 2:
            obs_covs = obs_covs[:, :, None]
        if obs_covs.ndim == 3:
            obs_covs = obs_covs[:, None, :, :]
    if obs is not None:
        if obs.ndim == 2:
            obs = obs[:, None, :]
        if obs.ndim == 3:  # (S,P,J) -> single species
            obs = obs[None, ...]
    if session_duration is not None and session_duration.ndim == 2:
        session_duration = session_duration[:, None, :]
    return site_covs, obs_covs, obs, session_duration


class OccupancyLikelihood:
    """Packed dataset resident in HBM + the fused logp/grad kernels for one of the accelerated models."""

    def __init__(
        self,
        model: str,
        site_covs,
        obs_covs,
        obs,
        session_duration=None,
        *,
        false_positives_constant: bool = False,
        false_positives_unoccupied: bool = False,
        max_abundance: int = 100,
        dtype: str = "float32",
        prior: bool = True,
        strict_math: bool = False,
        prior_beta: Tuple[float, float] = (0.0, 1.0),
        prior_alpha: Tuple[float, float] = (0.0, 1.0),
        prior_fp_beta: Tuple[float, float] = (2.0, 5.0),
        prior_fp_rate: float = 1.0,
        prior_mu_scale: float = 10.0,
        prior_sigma: Tuple[float, float] = (5.0, 1.0),
        site_random_effects: bool = False,
        obs_random_effects: bool = False,
        prior_site_re_sd_scale: float = 1.0,
        prior_obs_re_sd_scale: float = 1.0,
        device: int = 0,
        max_chains: int = 0,
    ):
        if model not in _lib.BL_MODEL:
            raise ValueError(f"unknown model {model!r}; the accelerated path covers {sorted(_lib.BL_MODEL)}")
        if dtype not in _DT:
            raise ValueError("dtype must be 'float32' or 'float64'")
        self._lib = _lib.load()
        self._h = C.c_void_p()
        site_covs, obs_covs, obs, session_duration = (
            _as_numpy(site_covs), _as_numpy(obs_covs), _as_numpy(obs), _as_numpy(session_duration))
        if site_covs is None or obs_covs is None or obs is None:
            raise ValueError("site_covs, obs_covs and obs are required")
        site_covs, obs_covs, obs, session_duration = ensure_period_dim(site_covs, obs_covs, obs, session_duration)
        # same assertions as occu.py:102-133
        if obs.ndim != 4:
            raise ValueError("obs must be of shape (n_species, n_sites, n_periods, n_replicates)")
        if site_covs.ndim != 2:
            raise ValueError("site_covs must be of shape (n_sites, n_site_covs)")
        if obs_covs.ndim != 4:
            raise ValueError("obs_covs must be of shape (n_sites, n_periods, n_replicates, n_obs_covs)")
        S, P, J, Ko = obs_covs.shape
        if site_covs.shape[0] != S:
            raise ValueError("site_covs and obs_covs must have the same number of sites")
        if obs.shape[1:] != (S, P, J):
            raise ValueError("obs must have shape (n_species, n_sites, n_periods, n_replicates)")
        n_species = obs.shape[0]
        if n_species > 1 and (site_random_effects or obs_random_effects):
            raise BiolithB200Error(-2, "unsupported", "random effects with n_species > 1 are outside the accelerated path")
        if session_duration is not None and model != "occu_cop":
            session_duration = None
        if session_duration is not None and session_duration.shape != (S, P, J):
            raise ValueError("session_duration must have shape (n_sites, n_periods, n_replicates)")
        if model == "occu_cs":
            # occu_cs.py:29-30: Normal(0, s) on mu0 / mu1 and Gamma(a, b) on sigma0 / sigma1 travel in the
            # prior_fp_* slots of bl_desc (include/biolith_b200.h)
            prior_fp_beta, prior_fp_rate = tuple(prior_sigma), float(prior_mu_scale)
        if site_random_effects or obs_random_effects:
            # occu.py:170-173: HalfNormal(scale) on the two sd's travels in the (otherwise unused) prior_fp_a / _b slots
            if model != "occu" or false_positives_constant or false_positives_unoccupied:
                raise BiolithB200Error(-2, "unsupported", "random effects are accelerated for occu without "
                                       "false-positive extras only")
            prior_fp_beta = (float(prior_site_re_sd_scale), float(prior_obs_re_sd_scale))
        code, npdt = _DT[dtype]
        # the reference casts to the compute dtype on ingestion (jnp.array, data.py:135-140)
        data_dt = np.float64 if any(a.dtype == np.float64 for a in (site_covs, obs_covs, obs)) else np.float32
        y = np.ascontiguousarray(obs, dtype=data_dt)
        X = np.ascontiguousarray(site_covs, dtype=data_dt)
        W = np.ascontiguousarray(obs_covs, dtype=data_dt)
        T = None if session_duration is None else np.ascontiguousarray(session_duration, dtype=data_dt)
        flags = (_lib.BL_FLAG_FP_CONSTANT if false_positives_constant else 0) | (
            _lib.BL_FLAG_FP_UNOCCUPIED if false_positives_unoccupied else 0) | (_lib.BL_FLAG_PRIOR if prior else 0) | (
            _lib.BL_FLAG_STRICT_MATH if strict_math else 0) | (_lib.BL_FLAG_SITE_RE if site_random_effects else 0) | (
            _lib.BL_FLAG_OBS_RE if obs_random_effects else 0)
        d = bl_desc(
            abi_version=_lib.BL_ABI_VERSION, model=_lib.BL_MODEL[model], dtype=code,
            data_dtype=_lib.BL_F64 if data_dt == np.float64 else _lib.BL_F32, flags=flags, device=device,
            n_sites=S, n_periods=P, n_replicates=J, n_site_covs=site_covs.shape[1], n_obs_covs=Ko, n_species=n_species,
            max_abundance=int(max_abundance), max_chains=int(max_chains), reserved0=0,
            prior_beta_loc=prior_beta[0], prior_beta_scale=prior_beta[1],
            prior_alpha_loc=prior_alpha[0], prior_alpha_scale=prior_alpha[1],
            prior_fp_a=prior_fp_beta[0], prior_fp_b=prior_fp_beta[1], prior_fp_rate=prior_fp_rate,
        )
        check(self._lib.bl_dataset_create(
            C.byref(d), y.ctypes.data, X.ctypes.data, W.ctypes.data,
            None if T is None else T.ctypes.data, C.byref(self._h)), "bl_dataset_create")
        info = bl_info()
        check(self._lib.bl_dataset_info(self._h, C.byref(info)), "bl_dataset_info")
        self.model, self.dtype, self.np_dtype, self.device = model, dtype, npdt, device
        self.shape = dict(n_sites=S, n_periods=P, n_replicates=J, n_site_covs=site_covs.shape[1], n_obs_covs=Ko)
        self.n_species = n_species
        self.site_random_effects, self.obs_random_effects = bool(site_random_effects), bool(obs_random_effects)
        self.theta_dim = info.theta_dim
        self.n_extras = info.n_extras
        self.packed_bytes = info.packed_bytes
        self.algorithmic_bytes = info.algorithmic_bytes
        self.n_masked = info.n_masked
        self.kernel_variant = info.kernel_variant
        self.fields_per_unit = info.fields_per_unit

    # -- lifetime --------------------------------------------------------------------------
    @property
    def handle(self) -> int:
        if not self._h:
            raise BiolithB200Error(-1, "handle", "dataset already destroyed")
        return self._h.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bl_dataset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the path --------------------------------------------------------------------------
    def logp_and_grad(self, theta) -> Tuple[np.ndarray, np.ndarray]:
        """theta (C, D) or (D,) host array -> (logp (C,), grad (C, D)); includes H2D/D2H."""
        th = np.ascontiguousarray(theta, dtype=self.np_dtype)
        single = th.ndim == 1
        if single:
            th = th[None, :]
        if th.ndim != 2 or th.shape[1] != self.theta_dim:
            raise ValueError(f"theta must have shape (C, {self.theta_dim}), got {th.shape}")
        n = th.shape[0]
        logp = np.empty(n, dtype=self.np_dtype)
        grad = np.empty((n, self.theta_dim), dtype=self.np_dtype)
        check(self._lib.bl_eval_host(self._h, th.ctypes.data, n, logp.ctypes.data, grad.ctypes.data), "bl_eval_host")
        return (logp[0], grad[0]) if single else (logp, grad)

    def eval_device(self, theta_ptr: int, n_chains: int, logp_ptr: int, grad_ptr: int, stream: int = 0):
        """Asynchronous evaluation on device pointers (what a jax.ffi custom call does)."""
        check(self._lib.bl_eval(self._h, theta_ptr, n_chains, logp_ptr, grad_ptr, stream), "bl_eval")

    def plan(self, n_chains: int) -> dict:
        """Which kernel evaluates a batch of ``n_chains`` chains and with what launch geometry (diagnostics / tests):
        ``kernel`` 0 = site-parallel engine, 5 = K1d, 6 = K2d, 7 = K1s (small batches), 1-4 = round-1 chain kernels."""
        k, gx, gy, bt = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self._lib.bl_plan_kernel(self._h, int(n_chains), C.byref(k), C.byref(gx), C.byref(gy), C.byref(bt)),
              "bl_plan_kernel")
        return dict(kernel=k.value, grid=(gx.value, gy.value), block_threads=bt.value)

    def eval_timed(self, theta_ptr, n_chains, logp_ptr, grad_ptr, stream=0, iters=10) -> float:
        ms = C.c_float()
        check(self._lib.bl_eval_timed(self._h, theta_ptr, n_chains, logp_ptr, grad_ptr, stream, iters,
                                      C.byref(ms)), "bl_eval_timed")
        return float(ms.value)

    def site_summary(self, theta_draws) -> dict:
        """Per-unit posterior summaries over a batch of draws, streamed on the GPU (no draws x sites
        array is ever formed).  theta_draws: (N, D).  Arrays are (n_sites, n_periods).

        ``psi_mean`` (occu_rn: ``abundance_mean``), ``occupancy_prob`` = mean P(z=1 | y) (occu_rn:
        ``abundance_posterior_mean`` = mean E[N | y]), ``lppd`` and ``p_waic`` pointwise on the
        marginalised unit, and their totals / ``waic`` on the deviance scale."""
        th = np.ascontiguousarray(theta_draws, dtype=self.np_dtype)
        if th.ndim != 2 or th.shape[1] != self.theta_dim:
            raise ValueError(f"theta_draws must have shape (N, {self.theta_dim})")
        s = self.shape
        U = s["n_sites"] * s["n_periods"]
        out = np.empty((4, U), dtype=np.float32)
        check(self._lib.bl_site_summary(self._h, th.ctypes.data, th.shape[0], out.ctypes.data), "bl_site_summary")
        out = out.reshape(4, s["n_sites"], s["n_periods"])
        rn = self.model in ("occu_rn", "nmixture")
        lppd, p_waic = out[2].astype(np.float64), out[3].astype(np.float64)
        return {
            ("abundance_mean" if rn else "psi_mean"): out[0],
            ("abundance_posterior_mean" if rn else "occupancy_prob"): out[1],
            "lppd": out[2], "p_waic": out[3],
            "lppd_total": float(lppd.sum()), "p_waic_total": float(p_waic.sum()),
            "waic": float(-2.0 * (lppd.sum() - p_waic.sum())),
        }

    def pointwise_loglik(self, theta_draws) -> dict:
        """Per-observation lppd / p_waic over a batch of draws with the definitions of the reference's
        ``lppd`` / ``waic`` (biolith/evaluation/lppd.py:51-61, waic.py:60-82), z integrated out per draw as in its
        closed form ``log_likelihood_manual`` (evaluation/log_likelihood.py:55-98).  Streamed on the GPU: arrays are
        (n_sites, n_periods, n_replicates), NaN where the observation is masked; totals run over valid observations."""
        th = np.ascontiguousarray(theta_draws, dtype=self.np_dtype)
        if th.ndim != 2 or th.shape[1] != self.theta_dim:
            raise ValueError(f"theta_draws must have shape (N, {self.theta_dim})")
        s = self.shape
        shp = (s["n_sites"], s["n_periods"], s["n_replicates"])
        lppd = np.empty(shp, dtype=np.float32)
        var = np.empty(shp, dtype=np.float32)
        check(self._lib.bl_obs_loglik(self._h, th.ctypes.data, th.shape[0], lppd.ctypes.data, var.ctypes.data),
              "bl_obs_loglik")
        tot, pw = float(np.nansum(lppd.astype(np.float64))), float(np.nansum(var.astype(np.float64)))
        return {"lppd": lppd, "p_waic": var, "lppd_total": tot, "p_waic_total": pw, "waic": -2.0 * (tot - pw)}

    def mask(self) -> np.ndarray:
        """(S, P, J) bool: which observations enter the likelihood (the bit-exact mask contract)."""
        s = self.shape
        shp = (s["n_sites"], s["n_periods"], s["n_replicates"])
        out = np.empty(shp if self.n_species == 1 else (self.n_species,) + shp, dtype=np.uint8)
        check(self._lib.bl_dataset_export_mask(self._h, out.ctypes.data), "bl_dataset_export_mask")
        return out.astype(bool)


class DeviceBuffer:
    """Tiny RAII wrapper over bl_device_malloc so that the host side needs no torch / cupy."""

    def __init__(self, nbytes: int, device: int = 0):
        self._lib = _lib.load()
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        check(self._lib.bl_device_malloc(device, self.nbytes, C.byref(self.ptr)), "bl_device_malloc")

    def upload(self, arr: np.ndarray, stream: int = 0):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(self._lib.bl_memcpy_h2d(self.ptr, arr.ctypes.data, arr.nbytes, stream), "bl_memcpy_h2d")
        check(self._lib.bl_stream_sync(stream), "bl_stream_sync")

    def download(self, shape, dtype, stream: int = 0) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(self._lib.bl_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, stream), "bl_memcpy_d2h")
        check(self._lib.bl_stream_sync(stream), "bl_stream_sync")
        return out

    def free(self):
        if self.ptr and self.ptr.value:
            self._lib.bl_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
