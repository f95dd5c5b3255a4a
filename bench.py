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


def host_threads():
    """Threads the CPU arm may use: the process's CPU affinity, NOT OMP_NUM_THREADS (torchrun exports
    OMP_NUM_THREADS=1 to every rank, which made the round-1 reference arm single-threaded at N >= 2)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


L2_NOTE = "flushed between timed steps (256 MB memset outside the per-step CUDA events)"


def build_config(args, model, X, W, chains, world, shard):
    """The workload description both arms print (same keys, same values: the unit is per chain-eval)."""
    return {"workload": args.workload, "likelihood": model, "strict_math": bool(args.strict_math),
            "sites": int(X.shape[0] * (world if shard == "sites" else 1)), "visits": int(W.shape[2]),
            "site_covs": int(X.shape[1]), "obs_covs": int(W.shape[3]), "chains_per_gpu": chains,
            "chains_total": chains * (world if shard == "chains" else 1), "sharding": shard,
            "theta": "U(-2,2) per chain (init_to_uniform)" if args.theta == "uniform"
            else "within 0.01 of the simulating truth (near the posterior mode)", "l2": L2_NOTE}


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


def run_reference(args, rank, world):
    """CPU arm: the C/OpenMP restatement of the reference's path (the reference itself needs
    jax+numpyro, which are neither installed nor installable on this image), all host threads."""
    if rank != 0:
        return
    from oracle import c_oracle

    model, X, W, y, chains, shard = make_data(args.workload, 0)
    if args.chains:
        chains = args.chains
    threads = host_threads()
    D = X.shape[1] + W.shape[3] + 2
    sample = args.cpu_chains
    th = np.random.default_rng(1).uniform(-2, 2, size=(sample, D))
    for _ in range(args.warmup):
        c_oracle.occu_logp_grad(th[:1], X, W, y, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.occu_logp_grad(th, X, W, y, nthreads=threads)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": build_config(args, model, X, W, chains, world, shard),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"each timed step = {sample} chain-evals over the full {X.shape[0]} x {W.shape[2]} "
                                   f"dataset (a bounded sample of the {chains}-chain step; the unit is per chain-eval), "
                                   f"{args.steps} steps, fp32, C/OpenMP restatement oracle/occu_oracle.c, {threads} threads "
                                   f"(CPU affinity; OMP_NUM_THREADS ignored)"},
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
    ap.add_argument("--nuts-warmup", type=int, default=0,
                    help="default: 1000 at N = 1 (the metric's own schedule, fit.py:22-23), 300 at N > 1")
    ap.add_argument("--nuts-samples", type=int, default=0, help="default: as --nuts-warmup")
    ap.add_argument("--no-nuts", action="store_true", help="skip the NUTS ESS/s section")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the configs[2]/[3] lines (N=1)")
    ap.add_argument("--no-site-sharded", action="store_true", help="skip the configs[4] block (N>1)")
    ap.add_argument("--site-sharded-timeout", type=float, default=300.0,
                    help="watchdog of the configs[4] block: the headline line is printed regardless")
    ap.add_argument("--strict-math", action="store_true", help="BL_FLAG_STRICT_MATH (libm expf/log1pf)")
    ap.add_argument("--theta", default="uniform", choices=["uniform", "mode"],
                    help="chain positions: U(-2,2) (init_to_uniform, default) or within 0.01 of the simulating truth "
                         "(SURVEY 8d asks for both; the kernels have no theta-dependent branches)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    n_ranks = int(os.environ.get("WORLD_SIZE", "1"))
    if args.nuts_warmup <= 0:
        args.nuts_warmup = 1000 if n_ranks == 1 else 300
    if args.nuts_samples <= 0:
        args.nuts_samples = 1000 if n_ranks == 1 else 300

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

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
    line = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        ms_launch = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_launch, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": build_config(args, model, X, W, chains, world, shard),
            "packed_dataset_mb": lk.packed_bytes / 1e6,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(theta.nbytes),
                    "d2h_bytes_per_step": int(theta.nbytes + chains * es),
                    "note": "bl_eval_host: pinned H2D theta + kernel + D2H logp,grad per step; dataset packed once per fit"},
            "gpu_launches": int(launches),
            **rooflines(lib, local_rank, args.workload, model, lk, X, W, chains, ms_launch, args),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if nuts is not None:
            line["nuts"] = nuts
        if world == 1 and args.workload == "occu_1m_x8_c1024" and not args.no_other_workloads:
            line["other_workloads"] = other_workloads(lib, local_rank, args)
        if world == 1 and model == "occu" and args.dtype == "float32" and not args.strict_math and not args.no_other_workloads:
            line["small_batch"] = small_batch(lib, lk, local_rank)
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
                                       "note": "BL_FLAG_STRICT_MATH: FMA-pipe exp2, libm log2f and IEEE division (no MUFU) in the K1d formulation (engine: BL_STRICT_ENGINE=1), same data and thetas"}
            except Exception as exc:  # noqa: BLE001
                line["strict_math"] = {"error": str(exc)[:200]}
        if not args.no_cpu_baseline and world == 1 and model in ("occu", "occu_cop", "occu_rn"):
            line["cpu_baseline"] = cpu_baseline(X, W, y, D, args.cpu_chains, model, make_data.session_duration)
    lk.close()
    if world > 1 and shard == "chains" and args.workload == "occu_1m_x8_c1024" and not args.no_site_sharded:
        # the one multi-GPU mode with a data-path collective (BASELINE configs[4]); measured in the same run so that
        # the driver's SCALE record carries it -- LAST and under a watchdog: whatever happens in here (a peer that
        # never arrives in a collective), rank 0 still prints the line it has already measured and every rank exits
        def give_up():
            if rank == 0:
                line["site_sharded"] = {"error": f"no result within {args.site_sharded_timeout} s (watchdog)"}
                print(json.dumps(_finite(line)), flush=True)
            os._exit(0)

        dog = threading.Timer(args.site_sharded_timeout, give_up)
        dog.daemon = True
        dog.start()
        try:
            site_sharded = run_site_sharded(args, dist, rank, world, local_rank)
        except Exception as exc:  # noqa: BLE001 - never fatal for the headline line
            site_sharded = {"error": str(exc)[:300]}
        dog.cancel()
        if rank == 0:
            line["site_sharded"] = site_sharded
    if rank == 0:
        print(json.dumps(_finite(line)), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# Per-(unit, chain) instruction counts of the shipped lane = chain kernels, from the committed ncu captures
# (profiles/; sm__inst_executed.sum and the MUFU share of it divided by units x chains of the profiled launch).
# They turn a measured time into an achieved issue / SFU rate; the peaks are measured by bl_pipe_peak.
PROFILED = {
    # workload: (warp-instructions per (unit, chain) x 32 lanes -> per lane, MUFU per (unit, chain), DRAM bytes / launch, source)
    # K1d (csrc/occu_signed.cu): 268.49 MB read + 7.31 MB written per launch (signed records: 176 MB packed)
    "occu_1m_x8_c1024": dict(instr=198.8, mufu=15.1, traffic=275_796_480, src="profiles/r02_occu_signed_final.txt"),
    "occu_cop_500k_x12": dict(instr=368.0, mufu=44.0, traffic=None, src="profiles/r01_occu_cop_chain_v1.txt"),
    # K2d (csrc/occu_rn2.cu): 41.80 MB read + 0.38 MB written per launch
    "occu_rn_200k_x10_k50": dict(instr=4816.3, mufu=152.0, traffic=42_176_512, src="profiles/r02_occu_rn2_final.txt"),
}
_PIPE_PEAKS = {}


def pipe_peaks(lib, device):
    """Measured MUFU lane-ops/s and warp-instruction issue/s of this GPU (csrc/microbench.cu), once per process."""
    if device not in _PIPE_PEAKS:
        from biolith_b200 import _lib

        out = {}
        for which, key in ((0, "mufu_per_s"), (1, "issue_per_s")):
            v, clk = C.c_double(), C.c_double()
            _lib.check(lib.bl_pipe_peak(device, which, C.byref(v), C.byref(clk)), "bl_pipe_peak")
            out[key] = float(v.value)
        _PIPE_PEAKS[device] = out
    return _PIPE_PEAKS[device]


def rooflines(lib, device, workload, model, lk, X, W, chains, ms_launch, args):
    """`roofline` = the unit that BINDS the dominant kernel.  For the chain-batched kernels that is instruction
    issue (and the SFU pipe next to it): a staged site tile is reused by every chain of the block, so physical
    DRAM traffic is ~1.6x the packed dataset per launch (0.4 % of HBM).  The SURVEY 8d figure -- C x B_eval
    algorithmic bytes per launch over the HBM peak -- is kept as `roofline_hbm_algorithmic`."""
    peak_hbm, peak_src = measured_peak_hbm()
    alg_bytes = lk.algorithmic_bytes * chains
    t = ms_launch * 1e-3
    units = float(lk.shape["n_sites"] * lk.shape["n_periods"])
    out = {}
    hbm = {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
           "frac": alg_bytes / t / 1e9 / peak_hbm, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": int(alg_bytes),
           "note": "SURVEY 8d definition: C x B_eval algorithmic bytes per launch (every chain's density is a full "
                   "pass over the data); with tile reuse across the chain batch frac > 1 is expected and is NOT a "
                   "physical bandwidth"}
    prof = PROFILED.get(workload)
    chain_kernel = (prof is not None and args.dtype == "float32" and not args.strict_math
                    and chains >= (64 if model == "occu_cs" else 32))
    if chain_kernel:
        pk = pipe_peaks(lib, device)
        winstr = prof["instr"] * units * chains  # warp-instructions x lanes... = thread-instructions
        issue_rate = winstr / 32.0 / t           # warp-instructions per second
        out["roofline"] = {
            "bound": "issue", "achieved": issue_rate / 1e9, "peak": pk["issue_per_s"] / 1e9, "unit": "G warp-instr/s",
            "frac": issue_rate / pk["issue_per_s"], "traffic": prof["traffic"],
            "instr_per_unit_chain": prof["instr"], "instr_source": prof["src"],
            "peak_source": "measured: bl_pipe_peak(issue), independent FFMA chains on all 4 schedulers of every SM",
            "note": "binding unit of the lane = chain kernel; `traffic` = DRAM bytes per launch from the ncu capture"}
        if prof["mufu"]:
            mufu_rate = prof["mufu"] * units * chains / t
            out["roofline_sfu"] = {"bound": "sfu", "achieved": mufu_rate / 1e12, "peak": pk["mufu_per_s"] / 1e12,
                                   "unit": "T MUFU/s", "frac": mufu_rate / pk["mufu_per_s"],
                                   "mufu_per_unit_chain": prof["mufu"],
                                   "peak_source": "measured: bl_pipe_peak(mufu), MUFU.EX2 chains"}
        out["roofline_hbm_algorithmic"] = hbm
    else:
        hbm["traffic"] = prof["traffic"] if prof else None
        out["roofline"] = hbm
    return out


def _time_eval(lib, lk, theta, device, steps, warmup=3, iters=1, flush=True):
    """ms per evaluation (CUDA events on the launch stream, L2 flushed between steps) for a resident theta."""
    import biolith_b200 as bb
    from biolith_b200 import _lib

    stream = C.c_void_p()
    _lib.check(lib.bl_stream_create(device, C.byref(stream)), "bl_stream_create")
    n = theta.shape[0]
    es = theta.dtype.itemsize
    d_th, d_lp, d_gr = (bb.DeviceBuffer(theta.nbytes, device), bb.DeviceBuffer(n * es, device),
                        bb.DeviceBuffer(theta.nbytes, device))
    d_th.upload(theta, stream)
    ms = []
    for i in range(warmup + steps):
        if flush:
            _lib.check(lib.bl_flush_l2(device, stream), "flush")
        v = lk.eval_timed(d_th.ptr, n, d_lp.ptr, d_gr.ptr, stream, iters=iters)
        if i >= warmup:
            ms.append(v)
    for b in (d_th, d_lp, d_gr):
        b.free()
    lib.bl_stream_destroy(stream)
    return float(np.mean(ms))


def small_batch(lib, lk, device):
    """The same dataset evaluated for 1 and 5 chains (5 = the reference's default num_chains, utils/fit.py:24): the
    regime that IS bound by HBM -- one pass over the packed dataset per evaluation, nothing to reuse a tile for --
    so its roofline is bytes / time against the measured copy bandwidth (L2 flushed between steps)."""
    out = {}
    peak, src = measured_peak_hbm()
    try:
        for c in (1, 5):
            theta = np.random.default_rng(2000 + c).uniform(-2, 2, size=(c, lk.theta_dim)).astype(np.float32)
            ms_flushed = _time_eval(lib, lk, theta, device, steps=20, warmup=5)
            ms = _time_eval(lib, lk, theta, device, steps=5, warmup=2, iters=200, flush=False)
            plan = lk.plan(c)
            gbs_alg = c * lk.algorithmic_bytes / (ms * 1e-3) / 1e9
            gbs_phys = lk.packed_bytes / (ms * 1e-3) / 1e9
            out[f"c{c}"] = {"chains": c, "us_per_eval": ms * 1e3, "us_per_eval_single_launch_after_l2_flush": ms_flushed * 1e3,
                            "timing": "CUDA events around 200 back-to-back launches; the packed dataset is larger than "
                                      "L2 (ncu: dram bytes read = packed bytes per launch, profiles/r02_occu_small_c*.txt)",
                            "value": c / (ms * 1e-3), "unit": UNIT,
                            "kernel": plan["kernel"], "grid": list(plan["grid"]), "block_threads": plan["block_threads"],
                            "roofline": {"bound": "hbm", "achieved": gbs_phys, "peak": peak, "unit": "GB/s",
                                         "frac": gbs_phys / peak, "traffic": lk.packed_bytes,
                                         "achieved_algorithmic": gbs_alg, "frac_algorithmic": gbs_alg / peak,
                                         "peak_source": src,
                                         "note": "achieved = packed dataset bytes (one pass, = ncu dram bytes read) / "
                                                 "time; achieved_algorithmic = chains x SURVEY 8d bytes / time"}}
    except Exception as exc:  # noqa: BLE001 - never fatal for the headline line
        out["error"] = str(exc)[:300]
    return out


def other_workloads(lib, device, args):
    """BASELINE configs[2] (occu_rn 200k x 10, K = 50) and configs[3] (occu_cop 500k x 12, NaN-masked) timed in the
    same run so that the driver's record carries them; same timing rules, fewer steps."""
    import biolith_b200 as bb

    out = {}
    for wl, chains in (("occu_rn_200k_x10_k50", 256), ("occu_cop_500k_x12", 1024)):
        try:
            model, X, W, y, _, _ = make_data(wl, 0)
            T = make_data.session_duration
            with bb.OccupancyLikelihood(model, X, W, y, T, device=device, max_chains=chains,
                                        **MODEL_KW.get(model, {})) as lk:
                theta = np.random.default_rng(1000).uniform(-2, 2, size=(chains, lk.theta_dim)).astype(np.float32)
                ms = _time_eval(lib, lk, theta, device, steps=5)
                sub = argparse.Namespace(**{**vars(args), "workload": wl})
                entry = {"ms_per_step": ms, "value": chains / (ms * 1e-3), "unit": UNIT,
                         "config": build_config(sub, model, X, W, chains, 1, "chains"),
                         **rooflines(lib, device, wl, model, lk, X, W, chains, ms, args)}
            out[wl] = entry
        except Exception as exc:  # noqa: BLE001 - never fatal for the headline line
            out[wl] = {"error": str(exc)[:300]}
    return out


def run_site_sharded(args, dist, rank, world, device):
    """BASELINE configs[4]: occu, sites sharded over the ranks, 256 chains, one exchange of the [C, 1+D] fp64 sums
    per evaluation.  Weak (2M sites per rank) and strong (a fixed 16M-site problem cut into `world` shards) sizes,
    both exchange modes, the exchange cost isolated (attached minus plain evaluation of the same shard), and the
    summed result checked against the fp64 C oracle evaluated on every rank's own shard."""
    import torch

    import biolith_b200 as bb
    from biolith_b200 import sharded
    from oracle import c_oracle

    lib = bb._lib.load()
    chains = 256
    out = {"chains": chains, "world": world}
    t_start = time.perf_counter()

    def stage(msg):  # progress on stderr: a watchdog exit then shows how far the block got
        if rank == 0:
            print(f"[site_sharded +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    sizes = {"weak_2m_sites_per_rank": 2_000_000}
    if 16_000_000 // world != 2_000_000:
        sizes[f"strong_16m_sites_over_{world}_ranks"] = 16_000_000 // world
    theta = np.random.default_rng(77).uniform(-2, 2, size=(chains, 10)).astype(np.float32)  # same on all ranks

    def tmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    for label, n_sites in sizes.items():
        model, kw, _, _ = WORKLOADS["occu_sites16m_c256"]
        WORKLOADS["_site_tmp"] = (model, {**kw, "n_sites": n_sites}, chains, "sites")
        stage(f"{label}: generating {n_sites} sites per rank")
        _, X, W, y, _, _ = make_data("_site_tmp", rank)
        entry = {"sites_per_rank": n_sites, "sites_total": n_sites * world}
        stage(f"{label}: local-only timing")
        with bb.OccupancyLikelihood(model, X, W, y, None, device=device, max_chains=chains) as plain:
            dist.barrier()
            entry["ms_per_eval_local_only"] = tmax(_time_eval(lib, plain, theta, device, steps=args.steps))
        ref = None
        for mode in ("p2p", "nccl"):
            stage(f"{label}: exchange mode {mode}")
            with bb.OccupancyLikelihood(model, X, W, y, None, device=device, max_chains=chains) as lk:
                sharded.attach_site_sharding(lk, dist, rank, world, chains, mode=mode)
                dist.barrier()
                ms = tmax(_time_eval(lib, lk, theta, device, steps=args.steps))
                lp, gr = lk.logp_and_grad(theta)
                if sharded.comm_error(lk):
                    raise RuntimeError("cross-GPU exchange timed out")
                got = np.concatenate([lp.astype(np.float64)[:, None], gr.astype(np.float64)], axis=1)
                g = torch.tensor(got, dtype=torch.float64, device="cuda")
                gmax, gmin = g.clone(), g.clone()
                dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
                if ref is None:
                    # fp64 C oracle on THIS rank's shard (4 chains, likelihood only), summed over ranks, priors once
                    k = 4
                    th64 = theta[:k].astype(np.float64)
                    threads = max(1, host_threads() // world)
                    olp, ogr = c_oracle.occu_logp_grad(th64, X, W, y, dtype=np.float64, prior=False, nthreads=threads)
                    t = torch.tensor(np.concatenate([olp[:, None], ogr], axis=1), dtype=torch.float64, device="cuda")
                    dist.all_reduce(t)
                    ref = t.cpu().numpy()
                    ref[:, 0] += (-0.5 * th64 ** 2 - 0.9189385332046727).sum(axis=1)
                    ref[:, 1:] -= th64
                err = np.abs(got[:ref.shape[0]] - ref) / np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
                err_lp = np.abs(got[:ref.shape[0], 0] - ref[:, 0]) / np.abs(ref[:, 0])
                entry[mode] = {
                    "ms_per_eval": ms, "chain_evals_per_s": chains / (ms * 1e-3),
                    "exchange_ms": ms - entry["ms_per_eval_local_only"],
                    "max_rel_err_logp_vs_fp64_c_oracle": float(err_lp.max()),
                    "max_rel_err_grad_vs_fp64_c_oracle": float(err[:, 1:].max()),
                    "bit_identical_across_ranks": bool(torch.equal(gmax, gmin)),
                }
                lib.bl_dataset_detach_comm(lk.handle)
            dist.barrier()
        out[label] = entry
        del X, W, y
    if world >= 4 and world % 2 == 0:
        # hybrid chains x sites grid (third row): site groups of 2 ranks share 256 chains and exchange
        # inside the group only; the world / 2 groups run different chains and never talk
        model, kw, _, _ = WORKLOADS["occu_sites16m_c256"]
        WORKLOADS["_site_tmp"] = (model, {**kw, "n_sites": 2_000_000}, chains, "sites")
        g, site_rank, _ = sharded.hybrid_layout(rank, world, 2)
        stage("hybrid: generating data")
        _, X, W, y, _, _ = make_data("_site_tmp", site_rank)
        stage("hybrid: attach + timing")
        th_g = np.random.default_rng(500 + g).uniform(-2, 2, size=(chains, 10)).astype(np.float32)
        entry = {"site_group_size": 2, "chain_groups": world // 2, "sites_per_rank": 2_000_000,
                 "chains_total": chains * (world // 2)}
        with bb.OccupancyLikelihood(model, X, W, y, None, device=device, max_chains=chains) as lk:
            _, _, sub = sharded.attach_hybrid(lk, dist, rank, world, 2, chains, mode="p2p")
            dist.barrier()
            ms = tmax(_time_eval(lib, lk, th_g, device, steps=args.steps))
            lp, gr = lk.logp_and_grad(th_g)
            k = 4
            th64 = th_g[:k].astype(np.float64)
            olp, ogr = c_oracle.occu_logp_grad(th64, X, W, y, dtype=np.float64, prior=False,
                                               nthreads=max(1, host_threads() // world))
            t = torch.tensor(np.concatenate([olp[:, None], ogr], axis=1), dtype=torch.float64, device="cuda")
            dist.all_reduce(t, group=sub._group)
            ref = t.cpu().numpy()
            ref[:, 0] += (-0.5 * th64 ** 2 - 0.9189385332046727).sum(axis=1)
            ref[:, 1:] -= th64
            got = np.concatenate([lp.astype(np.float64)[:k, None], gr.astype(np.float64)[:k]], axis=1)
            err = np.abs(got - ref) / np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
            worst = torch.tensor([float(err.max())], dtype=torch.float64, device="cuda")
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            entry.update({"ms_per_eval": ms, "chain_evals_per_s": entry["chains_total"] / (ms * 1e-3),
                          "max_rel_err_vs_fp64_c_oracle_all_groups": float(worst[0]), "exchange": "p2p"})
            lib.bl_dataset_detach_comm(lk.handle)
        dist.barrier()
        out[f"hybrid_{world // 2}_chain_groups_x_2_site_shards"] = entry
    WORKLOADS.pop("_site_tmp", None)
    out["note"] = ("exchange = [C, 1+D] fp64 sums (22 KB): 'p2p' is ONE fused kernel over CUDA-IPC peer memory "
                   "(publish, flag, rank-ordered sum, priors), 'nccl' is ncclAllReduce + finalize; oracle check = fp64 "
                   "C/OpenMP restatement on every rank's own shard, summed with an fp64 all-reduce")
    return out if rank == 0 else None


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
                 rows=np.array([r["rows_evaluated"]]),
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
        # leapfrogs used by the chains / chain-evaluations spent (finished chains are compacted away in whole warps)
        "useful_eval_frac": float(leaps.sum() / max(float(np.sum(g["rows"])), 1.0)) if shard == "chains"
        else float(leaps.sum() / max(float(np.max(g["rows"])), 1.0)),
        "lockstep_eval_frac": float(leaps.sum() / (x.shape[0] * max(int(np.max(g["steps"])), 1))) if shard == "chains"
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


def cpu_baseline(X, W, y, D, sample, model="occu", T=None):
    from oracle import c_oracle

    threads = host_threads()
    th = np.random.default_rng(1).uniform(-2, 2, size=(sample, D))
    if model == "occu_cop":  # config 4: the C restatement computes in double (clamp constants of fp32)
        def run(t):
            return c_oracle.occu_cop_logp_grad(t, X, W, y, T, fp_constant=True, nthreads=threads)
        arith = "fp64-arithmetic"
    elif model == "occu_rn":  # config 3
        def run(t):
            return c_oracle.occu_rn_logp_grad(t, X, W, y, nthreads=threads, **MODEL_KW["occu_rn"])
        arith = "fp64-arithmetic"
    else:
        def run(t):
            return c_oracle.occu_logp_grad(t, X, W, y, nthreads=threads)
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
