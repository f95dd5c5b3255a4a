#!/usr/bin/env python
"""bench.py -- logp+grad evals/sec of the occupancy hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one batched evaluation of log-density + gradient for C chains over the whole
synthetic dataset (what NUTS does at one leapfrog step of every chain).  Workload (N=1 default):
BASELINE.json configs[1] -- occu, 5 site + 3 obs covariates, 1M sites x 8 visits, 1024 chains.
For N>1 (torchrun, one rank per GPU) chains are sharded: every rank holds the dataset and its own
1024 chains (weak scaling, no data-path collective); `--workload occu_sites16m` runs the
site-sharded configs[4] (2M sites per rank, per-eval allreduce of the [C, 1+D] sums).

Prints ONE JSON line (rank 0).  `value` = chain-evals/s with theta resident in HBM, CUDA-event
timed per step on the launch stream; `e2e` = the same through the host-buffer C-ABI call
(bl_eval_host: pinned H2D of theta, kernel, D2H of logp+grad every step); `roofline`,
`cpu_baseline`, `clocks`, `gpu_launches` as the harness contract asks.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, simulate kwargs, chains per GPU, shard mode)
    "occu_1m_x8_c1024": ("occu", dict(n_site_covs=5, n_obs_covs=3, n_sites=1_000_000, deployment_days_per_site=56),
                         1024, "chains"),
    "occu_sites16m_c256": ("occu", dict(n_site_covs=5, n_obs_covs=3, n_sites=2_000_000, deployment_days_per_site=56),
                           256, "sites"),
    "occu_small": ("occu", dict(n_site_covs=5, n_obs_covs=3, n_sites=20_000, deployment_days_per_site=56), 256,
                   "chains"),
    # BASELINE.json configs[2] / configs[3] (parity-test shapes; timed here for DESIGN.md, not the headline)
    "occu_rn_200k_x10_k50": ("occu_rn", dict(n_site_covs=5, n_obs_covs=3, n_sites=200_000,
                                             deployment_days_per_site=70), 256, "chains"),
    "occu_cop_500k_x12": ("occu_cop", dict(n_site_covs=5, n_obs_covs=3, n_sites=500_000,
                                           deployment_days_per_site=84, simulate_missing=True), 1024, "chains"),
    # SURVEY 8 row f4 siblings (site-parallel engine only), timed for DESIGN.md
    "nmixture_200k_x10": ("nmixture", dict(n_site_covs=5, n_obs_covs=3, n_sites=200_000,
                                           deployment_days_per_site=70), 256, "chains"),
    "occu_cs_500k_x10": ("occu_cs", dict(n_site_covs=5, n_obs_covs=3, n_sites=500_000,
                                         deployment_days_per_site=70), 256, "chains"),
}
MODEL_KW = {"occu_rn": dict(max_abundance=50), "occu_cop": dict(false_positives_constant=True)}
METRIC = "logp+grad evals/sec (occu, 1M sites x 8 visits, chain-batched); NUTS ESS/sec of the same run under 'nuts'"
UNIT = "chain-evals/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_data(workload, seed):
    from biolith_b200.simulate import simulate_occupancy

    model, kw, chains, shard = WORKLOADS[workload]
    if shard == "sites" and seed != 0:
        # every site shard shares ONE truth (that of seed 0); the seed only varies covariates / latents.
        # (Per-shard truths would make the pooled 16M-site posterior a product of 8 conflicting, extremely
        # sharp likelihoods whose local modes trap chains: measured r-hat 13-24 on 8 GPUs.)
        _, true0 = simulate_occupancy(model, random_seed=0, **{**kw, "n_sites": 2000})
        data, _ = simulate_occupancy(model, random_seed=seed, beta=true0["beta"], alpha=true0["alpha"], **kw)
    else:
        data, true0 = simulate_occupancy(model, random_seed=seed, **kw)
    make_data.true_theta = np.concatenate([true0["beta"][0], true0["alpha"][0]])
    X = data["site_covs"].astype(np.float32)   # the reference ingests as fp32 (utils/data.py:135-140)
    W = data["obs_covs"].astype(np.float32)
    y = data["obs"].astype(np.float32)
    T = data.get("session_duration")
    make_data.session_duration = None if T is None else T.astype(np.float32)
    return model, X, W, y, chains, shard


def run_reference(args, rank):
    """CPU arm: the C/OpenMP restatement of the reference's path (the reference itself needs
    jax+numpyro, which are neither installed nor installable on this image), all host threads."""
    if rank != 0:
        return
    from oracle import c_oracle

    model, X, W, y, chains, shard = make_data(args.workload, 0)
    threads = c_oracle.max_threads()
    D = X.shape[1] + W.shape[3] + 2
    sample = args.cpu_chains
    th = np.random.default_rng(1).uniform(-2, 2, size=(sample, D))
    for _ in range(args.warmup):
        c_oracle.occu_logp_grad(th[:1], X, W, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.occu_logp_grad(th, X, W, y)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sites": int(X.shape[0]), "visits": int(W.shape[2]),
                   "site_covs": int(X.shape[1]), "obs_covs": int(W.shape[3]),
                   "note": "each step = %d chain-evals over the full dataset (bounded sample of the 1024-chain step)" % sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} chain-evals x {args.steps} steps, full 1M x 8 dataset, fp32, OpenMP {threads} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="occu_1m_x8_c1024", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--cpu-chains", type=int, default=8, help="chain-evals per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "p2p"], help="site-sharded cross-rank sum")
    ap.add_argument("--nuts-warmup", type=int, default=200)
    ap.add_argument("--nuts-samples", type=int, default=100)
    ap.add_argument("--no-nuts", action="store_true", help="skip the NUTS ESS/s section")
    ap.add_argument("--strict-math", action="store_true", help="BL_FLAG_STRICT_MATH (libm expf/log1pf)")
    ap.add_argument("--theta", default="uniform", choices=["uniform", "mode"],
                    help="chain positions: U(-2,2) (init_to_uniform, default) or within 0.01 of the simulating truth "
                         "(SURVEY 8d asks for both; the kernels have no theta-dependent branches)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import biolith_b200 as bb
    from biolith_b200 import _lib

    lib = _lib.load()
    model, X, W, y, chains, shard = make_data(args.workload, 0 if shard_is_chains(args.workload) else rank)
    if args.chains:
        chains = args.chains
    if model == "nmixture":  # truncation point: comfortably above the largest count (nmixture.py:24)
        MODEL_KW["nmixture"] = dict(max_abundance=int(np.nanmax(y)) + 10)
    comm = None
    lk = bb.OccupancyLikelihood(model, X, W, y, make_data.session_duration, dtype=args.dtype, device=local_rank,
                                max_chains=chains, strict_math=args.strict_math, **MODEL_KW.get(model, {}))
    if shard == "sites" and world > 1:
        from biolith_b200 import sharded

        comm = sharded.attach_site_sharding(lk, dist, rank, world, chains, mode=args.exchange)
    D = lk.theta_dim
    npdt = lk.np_dtype
    es = np.dtype(npdt).itemsize
    # chain-sharded: every rank owns different chains; site-sharded: all ranks share the chains
    th_seed = 1000 + (rank if shard == "chains" else 0)
    theta = np.random.default_rng(th_seed).uniform(-2, 2, size=(chains, D)).astype(npdt)
    if args.theta == "mode":
        t0 = np.zeros(D)
        t0[: make_data.true_theta.size] = make_data.true_theta
        if model == "occu_cs":
            t0[-4:] = [0.0, np.log(10.0), np.log(10.0), np.log(5.0)]
        elif D > make_data.true_theta.size:
            t0[make_data.true_theta.size:] = np.log(0.1)
        theta = (t0 + 0.01 * np.random.default_rng(th_seed).standard_normal((chains, D))).astype(npdt)
    stream = C.c_void_p()
    _lib.check(lib.bl_stream_create(local_rank, C.byref(stream)), "bl_stream_create")
    d_theta = bb.DeviceBuffer(theta.nbytes, local_rank)
    d_logp = bb.DeviceBuffer(chains * es, local_rank)
    d_grad = bb.DeviceBuffer(theta.nbytes, local_rank)
    d_theta.upload(theta, stream)

    def barrier():
        _lib.check(lib.bl_stream_sync(stream), "sync")
        if dist is not None:
            dist.barrier()

    def step_timed():
        _lib.check(lib.bl_flush_l2(local_rank, stream), "flush")  # outside the per-step events
        return lk.eval_timed(d_theta.ptr, chains, d_logp.ptr, d_grad.ptr, stream, iters=1)

    for _ in range(args.warmup):
        step_timed()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.bl_launch_count()
    t_wall0 = time.perf_counter()
    ms = [step_timed() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.bl_launch_count() - launches0
    clocks = sampler.stop()
    total_ms = float(np.sum(ms))

    # e2e: host buffers through the public call, H2D + D2H inside the timed region
    for _ in range(2):
        lk.logp_and_grad(theta)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lp_host, gr_host = lk.logp_and_grad(theta)
    e2e_s = time.perf_counter() - t0
    barrier()

    if dist is not None:
        import torch

        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln[0])
    chains_total = chains * (world if shard == "chains" else 1)
    units_sites = X.shape[0] * (world if shard == "sites" else 1)
    value = chains_total * args.steps / (total_ms * 1e-3)
    e2e_value = chains_total * args.steps / e2e_s

    exchange_check = None
    if comm is not None:
        # self-check of the cross-rank exchange: the fused / NCCL result must equal the fp64 sum of every
        # rank's own shard evaluated on a plain (unattached) handle, and be bit-identical on all ranks
        import torch

        with bb.OccupancyLikelihood(model, X, W, y, make_data.session_duration, dtype=args.dtype, device=local_rank,
                                    prior=False, **MODEL_KW.get(model, {})) as loc:
            lp_loc, gr_loc = loc.logp_and_grad(theta[:64])
        t = torch.tensor(np.concatenate([lp_loc.astype(np.float64)[:, None], gr_loc.astype(np.float64)], axis=1),
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        ref = t.cpu().numpy()
        th64 = theta[:64].astype(np.float64)
        ref[:, 0] += (-0.5 * th64 ** 2 - 0.9189385332046727).sum(axis=1)  # the handle under test adds N(0,1) priors
        ref[:, 1:] -= th64
        got = np.concatenate([lp_host.astype(np.float64)[:64, None], gr_host.astype(np.float64)[:64]], axis=1)
        err = np.abs(got - ref) / np.maximum(np.abs(ref).max(axis=0, keepdims=True), 1.0)
        g = torch.tensor(got, dtype=torch.float64, device="cuda")
        gmax, gmin = g.clone(), g.clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
        exchange_check = {"max_rel_err_vs_sum_of_shards": float(err.max()),
                          "bit_identical_across_ranks": bool(torch.equal(gmax, gmin)), "mode": args.exchange}

    nuts = None
    if not args.no_nuts:
        nuts = run_nuts(args, lk, chains, rank, world, shard, dist)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        alg_bytes = lk.algorithmic_bytes * chains  # C x B_eval per launch (per GPU)
        ms_launch = total_ms / args.steps
        achieved = alg_bytes / (ms_launch * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_launch, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "likelihood": model, "strict_math": bool(args.strict_math), "sites": int(units_sites),
                       "visits": int(W.shape[2]), "site_covs": int(X.shape[1]), "obs_covs": int(W.shape[3]),
                       "chains_per_gpu": chains, "chains_total": chains_total, "sharding": shard,
                       "theta": "U(-2,2) per chain (init_to_uniform)" if args.theta == "uniform"
                       else "within 0.01 of the simulating truth (near the posterior mode)",
                       "l2": "flushed between timed steps (256 MB memset outside the per-step CUDA events); "
                             "packed dataset %.0f MB" % (lk.packed_bytes / 1e6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(theta.nbytes),
                    "d2h_bytes_per_step": int(theta.nbytes + chains * es),
                    "note": "bl_eval_host: pinned H2D theta + kernel + D2H logp,grad per step; dataset packed once per fit"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": PROFILED_TRAFFIC.get(args.workload),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "algorithmic bytes = C x B_eval (SURVEY 8d): every chain's density is a full pass "
                                 "over the data; tile reuse across the chain batch makes the kernel FP32/SFU-issue "
                                 "bound, so frac > 1 is expected (see DESIGN.md, profiles/)"},
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if model == "occu" and lk.kernel_variant == 1 and chains >= 64 and args.dtype == "float32" and not args.strict_math:
            # the unit that actually binds the chain kernel (profiles/): 3 MUFU (ex2, lg2, rcp) per logistic
            # term, (J + 2) terms per (site, chain); B200 SFU = 16 lanes / SM / clock
            mufu = 3.0 * (W.shape[2] + 2) * X.shape[0] * chains
            sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
            peak_mufu = 148 * 16 * sm_clock
            line["roofline_sfu"] = {"bound": "sfu", "achieved": mufu / (ms_launch * 1e-3) / 1e12,
                                    "peak": peak_mufu / 1e12, "unit": "T MUFU/s", "frac": mufu / (ms_launch * 1e-3) / peak_mufu,
                                    "note": "secondary roofline: the kernel is issue/SFU bound, not HBM bound"}
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if nuts is not None:
            line["nuts"] = nuts
        if world == 1 and model == "occu" and args.dtype == "float32" and not args.strict_math:
            # the same evaluation with libm expf / log1pf / IEEE division (BL_FLAG_STRICT_MATH), for the record:
            # the default kernels use bounded-error SFU forms (DESIGN.md "Numerics"); never fatal for the line
            try:
                with bb.OccupancyLikelihood(model, X, W, y, None, dtype=args.dtype, device=local_rank,
                                            max_chains=chains, strict_math=True) as strict:
                    strict.eval_timed(d_theta.ptr, chains, d_logp.ptr, d_grad.ptr, stream, iters=1)
                    ms_s = min(strict.eval_timed(d_theta.ptr, chains, d_logp.ptr, d_grad.ptr, stream, iters=2)
                               for _ in range(2))
                line["strict_math"] = {"ms_per_step": ms_s, "value": chains / (ms_s * 1e-3), "unit": UNIT,
                                       "note": "libm-accurate fp32 math in the site-parallel engine, same data and thetas"}
            except Exception as exc:  # noqa: BLE001
                line["strict_math"] = {"error": str(exc)[:200]}
        if not args.no_cpu_baseline and world == 1 and model in ("occu", "occu_cop", "occu_rn"):
            line["cpu_baseline"] = cpu_baseline(X, W, y, D, args.cpu_chains, model, make_data.session_duration)
        print(json.dumps(_finite(line)), flush=True)
    lk.close()
    if dist is not None:
        dist.destroy_process_group()


def run_nuts(args, lk, chains, rank, world, shard, dist):
    """NUTS ESS/sec: device-resident chain-batched NUTS (numpyro defaults: target 0.8, depth 10, diag mass),
    init U(-2,2); ESS by the numpyro.diagnostics definition over all chains of all ranks."""
    import biolith_b200 as bb
    from biolith_b200 import diagnostics as dg
    from biolith_b200 import sharded

    seed = 11 + (rank if shard == "chains" else 0)
    s = bb.NutsSampler(lk, chains, args.nuts_warmup, args.nuts_samples, seed=seed)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    ok = s.run(timeout=600)
    wall = time.perf_counter() - t0
    r = s.results()
    s.close()
    local = dict(samples=r["samples"], leapfrogs=r["leapfrogs"], warmup_leapfrogs=r["warmup_leapfrogs"],
                 num_steps=r["num_steps"], accept_prob=r["accept_prob"], diverging=r["diverging"],
                 step_size=r["step_size"], wall=np.array([wall]), steps=np.array([r["global_steps"]]),
                 ok=np.array([ok]))
    if dist is not None and shard == "chains":
        g = sharded.gather_chain_results(dist, rank, world, local)
    else:
        g = local
    if rank != 0:
        return None
    x = g["samples"].astype(np.float64)
    wall = float(np.max(g["wall"]))
    ne = dg.effective_sample_size(x)
    rhat = dg.split_gelman_rubin(x)
    leaps, wleaps = g["leapfrogs"].astype(np.float64), g["warmup_leapfrogs"].astype(np.float64)
    frac_sampling = float((leaps - wleaps).sum() / leaps.sum())
    return {
        "ess_per_sec_min": float(ne.min() / wall), "ess_per_sec_median": float(np.median(ne) / wall),
        "ess_per_sec_min_sampling_phase": float(ne.min() / (wall * frac_sampling)),
        "ess_min": float(ne.min()), "ess_median": float(np.median(ne)), "r_hat_max": float(rhat.max()),
        "chains": int(x.shape[0]), "num_warmup": args.nuts_warmup, "num_samples": args.nuts_samples,
        "wall_s": wall, "complete": bool(np.all(g["ok"])), "global_steps": int(np.max(g["steps"])),
        "ms_per_global_step": 1e3 * wall / max(int(np.max(g["steps"])), 1),
        "leapfrogs_per_chain_mean": float(leaps.mean()), "leapfrogs_per_draw": float(g["num_steps"].mean()),
        "useful_eval_frac": float(leaps.sum() / (x.shape[0] * max(int(np.max(g["steps"])), 1))) if shard == "chains"
        else float(leaps.mean() / max(int(np.max(g["steps"])), 1)),
        "accept_prob_mean": float(g["accept_prob"].mean()), "divergence_frac": float(g["diverging"].mean()),
        "step_size_median": float(np.median(g["step_size"])),
        "note": "wall includes warm-up; ESS = numpyro.diagnostics.effective_sample_size over all chains; "
                "min/median over the %d parameters" % x.shape[2],
    }


def _finite(o):
    """strict JSON: non-finite floats (e.g. ESS of a chain that never moved) become null"""
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    if isinstance(o, float) and not np.isfinite(o):
        return None
    return o


def shard_is_chains(workload):
    return WORKLOADS[workload][3] == "chains"


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the eval kernel (ncu --set full), by workload
PROFILED_TRAFFIC = {
    # profiles/r01_occu_chain_v6.txt (final round-1 binary): 227.34 MB read + 9.17 MB written per launch
    # (packed dataset: 128 MB; the 4 chain chunks re-read tiles mostly from L2)
    "occu_1m_x8_c1024": 236_505_600,
}


def cpu_baseline(X, W, y, D, sample, model="occu", T=None):
    from oracle import c_oracle

    threads = c_oracle.max_threads()
    th = np.random.default_rng(1).uniform(-2, 2, size=(sample, D))
    if model == "occu_cop":  # config 4: the C restatement computes in double (clamp constants of fp32)
        def run(t):
            return c_oracle.occu_cop_logp_grad(t, X, W, y, T, fp_constant=True)
        arith = "fp64-arithmetic"
    elif model == "occu_rn":  # config 3
        def run(t):
            return c_oracle.occu_rn_logp_grad(t, X, W, y, **MODEL_KW["occu_rn"])
        arith = "fp64-arithmetic"
    else:
        def run(t):
            return c_oracle.occu_logp_grad(t, X, W, y)
        arith = "fp32"
    run(th[:1])
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < 2.0 and reps < 50):
        run(th)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": sample * reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample * reps} chain-evals of the full dataset ({X.shape[0]} sites), {arith} C/OpenMP "
                      f"restatement (oracle/occu_oracle.c), {threads} threads, {dt:.1f} s wall"}


if __name__ == "__main__":
    main()
