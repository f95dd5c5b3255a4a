"""Host handle of the device-resident chain-batched NUTS sampler (bl_nuts_* in the C ABI)."""

from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np

from . import _lib
from ._lib import bl_nuts_config, check
from .likelihood import OccupancyLikelihood


class NutsSampler:
    """All chains advance one leapfrog per global step on the GPU; numpyro's NUTS defaults."""

    def __init__(self, likelihood: OccupancyLikelihood, num_chains: int, num_warmup: int = 1000,
                 num_samples: int = 1000, *, seed: int = 0, init_params: Optional[np.ndarray] = None,
                 init_radius: float = 2.0, max_tree_depth: int = 10, target_accept_prob: float = 0.8,
                 step_size: float = 1.0, adapt_step_size: bool = True, adapt_mass_matrix: bool = True,
                 find_heuristic_step_size: bool = False, validate_init: bool = True):
        self._lib = _lib.load()
        self.lk = likelihood
        self.num_chains, self.num_warmup, self.num_samples = int(num_chains), int(num_warmup), int(num_samples)
        D = likelihood.theta_dim
        drawn = init_params is None
        if drawn:  # init_to_uniform(radius=2), fit.py:93
            rng = np.random.default_rng(seed)
            init_params = rng.uniform(-init_radius, init_radius, size=(num_chains, D))
        th0 = np.ascontiguousarray(init_params, dtype=likelihood.np_dtype)
        if th0.shape != (num_chains, D):
            raise ValueError(f"init_params must have shape ({num_chains}, {D})")
        # numpyro's find_valid_initial_params: a start whose log-density or gradient is not finite is redrawn
        # (up to 100 times); a chain started there would reject every proposal and return identical draws
        if validate_init:
            for attempt in range(101):
                lp, gr = likelihood.logp_and_grad(th0)
                bad = ~(np.isfinite(lp) & np.isfinite(gr).all(axis=1))
                if not bad.any():
                    break
                if not drawn or attempt == 100:
                    raise _lib.BiolithB200Error(
                        -1, "NutsSampler", f"log-density or gradient is not finite at the initial position of chain(s) "
                        f"{np.flatnonzero(bad)[:8].tolist()}")
                th0[bad] = rng.uniform(-init_radius, init_radius, size=(int(bad.sum()), D))
        cfg = bl_nuts_config(
            n_chains=num_chains, num_warmup=num_warmup, num_samples=num_samples, max_tree_depth=max_tree_depth,
            adapt_step_size=int(adapt_step_size), adapt_mass_matrix=int(adapt_mass_matrix),
            find_heuristic_step_size=int(find_heuristic_step_size), reserved0=0, seed=seed,
            target_accept_prob=target_accept_prob, init_step_size=step_size, max_delta_energy=1000.0)
        self._h = C.c_void_p()
        check(self._lib.bl_nuts_create(likelihood.handle, C.byref(cfg), th0.ctypes.data, C.byref(self._h)),
              "bl_nuts_create")
        self.steps = 0
        self.wall_s = 0.0

    def run(self, max_steps: Optional[int] = None, poll_every: int = 64, timeout: Optional[float] = None):
        """Run until every chain has its draws (or max_steps / timeout).  Returns True if complete."""
        budget = max_steps if max_steps is not None else 1 << 62
        t0 = time.perf_counter()
        done = C.c_int32(0)
        steps = C.c_int64(0)
        chunk = max(poll_every, 1) * 16
        while budget > 0:
            n = min(chunk, budget)
            check(self._lib.bl_nuts_run(self._h, n, poll_every, C.byref(steps), C.byref(done)), "bl_nuts_run")
            budget -= n
            # site-sharded handles: a peer that timed out in the fused exchange leaves stale sums behind
            err = C.c_int32(0)
            check(self._lib.bl_dataset_comm_error(self.lk.handle, C.byref(err)), "bl_dataset_comm_error")
            if err.value:
                raise _lib.BiolithB200Error(-4, "bl_nuts_run", "a peer rank timed out in the cross-GPU exchange; "
                                            "the sampler state is not trustworthy")
            if done.value >= self.num_chains:
                break
            if timeout is not None and time.perf_counter() - t0 > timeout:
                break
        self.steps = steps.value
        self.wall_s += time.perf_counter() - t0
        return done.value >= self.num_chains

    def results(self) -> dict:
        Cn, N, D = self.num_chains, self.num_samples, self.lk.theta_dim
        samples = np.empty((N, Cn, D), np.float32)
        acc = np.empty((N, Cn), np.float32)
        nsteps = np.empty((N, Cn), np.int32)
        div = np.empty((N, Cn), np.uint8)
        pe = np.empty((N, Cn), np.float64)
        eps = np.empty(Cn, np.float64)
        imm = np.empty((D, Cn), np.float64)
        leaps = np.empty(Cn, np.int32)
        wleaps = np.empty(Cn, np.int32)
        saved = np.empty(Cn, np.int32)
        check(self._lib.bl_nuts_get(
            self._h, samples.ctypes.data, acc.ctypes.data, nsteps.ctypes.data, div.ctypes.data, pe.ctypes.data,
            eps.ctypes.data, imm.ctypes.data, leaps.ctypes.data, wleaps.ctypes.data, saved.ctypes.data),
            "bl_nuts_get")
        rows = C.c_int64(0)
        check(self._lib.bl_nuts_rows_evaluated(self._h, C.byref(rows)), "bl_nuts_rows_evaluated")
        return dict(
            rows_evaluated=int(rows.value),
            samples=samples.transpose(1, 0, 2), accept_prob=acc.T, num_steps=nsteps.T, diverging=div.T.astype(bool),
            potential_energy=pe.T, step_size=eps, inverse_mass_matrix=imm.T, leapfrogs=leaps,
            warmup_leapfrogs=wleaps, n_saved=saved, global_steps=self.steps, wall_s=self.wall_s)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bl_nuts_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
