"""GPU, needs >= 2 devices (skipped on a 1-GPU box): site-sharded evaluation over 2 ranks must equal
the single-handle evaluation of the concatenated dataset, bit-identically on both ranks, in both
exchange modes (NCCL allreduce / fused CUDA-IPC kernel); device NUTS stays in lock-step."""

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    """A free port BELOW the ephemeral range: a port handed out by bind(0) can be taken by another process's outgoing
    connection (e.g. an NCCL bootstrap socket) before the rendezvous store listens on it (seen once: EADDRINUSE)."""
    import random

    rng = random.Random(os.getpid() * 7919 + int.from_bytes(os.urandom(4), "little"))
    for _ in range(200):
        port = rng.randrange(21000, 31000)
        s = socket.socket()
        try:
            s.bind(("127.0.0.1", port))
            return port
        except OSError:
            continue
        finally:
            s.close()
    raise RuntimeError("no free port")


def _n_gpus():
    import ctypes as C

    from biolith_b200 import _lib

    n = C.c_int(0)
    rc = _lib.load().bl_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import biolith_b200 as bb
    from biolith_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=50_001,
                                        deployment_days_per_site=56, simulate_missing=True, random_seed=3)
        X, W, y, _ = sharded.shard_data(data["site_covs"], data["obs_covs"], data["obs"], None, rank, world)
        th = np.random.default_rng(0).uniform(-2, 2, size=(96, 10)).astype(np.float32)
        with bb.OccupancyLikelihood("occu", X, W, y, device=rank, max_chains=96) as lk:
            sharded.attach_site_sharding(lk, dist, rank, world, 96, mode=mode)
            lp, gr = lk.logp_and_grad(th)
            lp2, gr2 = lk.logp_and_grad(th[:5])  # site-parallel engine path, same comm
            s = bb.NutsSampler(lk, 96, 30, 20, seed=5)
            ok = s.run(timeout=120)
            res = s.results()
            s.close()
            err = sharded.comm_error(lk)
        ref = None
        if rank == 0:
            with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], device=0) as full:
                ref = full.logp_and_grad(th)
        q.put((rank, lp, gr, lp2, gr2, ok, res["samples"], err, ref))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_site_sharded_ranks(mode, world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    _, lp0, gr0, lpa0, gra0, ok0, s0, e0, ref = res[0]
    assert e0 == 0 and ok0
    for _, lp1, gr1, lpa1, gra1, ok1, s1, e1, _ in res[1:]:
        assert e1 == 0 and ok1
        assert np.array_equal(lp0, lp1) and np.array_equal(gr0, gr1), "ranks disagree bitwise"
        assert np.array_equal(lpa0, lpa1) and np.array_equal(gra0, gra1)
        assert np.array_equal(s0, s1), "site-sharded NUTS chains diverged between ranks"
    np.testing.assert_allclose(lp0, ref[0], rtol=2e-6)
    np.testing.assert_allclose(gr0, ref[1], rtol=1e-5, atol=1e-5 * np.abs(ref[1]).max())
    np.testing.assert_allclose(lpa0, ref[0][:5], rtol=2e-6)


def _hybrid_worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import biolith_b200 as bb
    from biolith_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g, site_rank, _ = sharded.hybrid_layout(rank, world, 2)
        # every chain group works on the SAME dataset here (so that one reference serves all) with its OWN chains
        data, _ = bb.simulate_occupancy("occu", n_site_covs=5, n_obs_covs=3, n_sites=40_003,
                                        deployment_days_per_site=56, simulate_missing=True, random_seed=3)
        X, W, y, _ = sharded.shard_data(data["site_covs"], data["obs_covs"], data["obs"], None, site_rank, 2)
        th = np.random.default_rng(100 + g).uniform(-2, 2, size=(64, 10)).astype(np.float32)
        with bb.OccupancyLikelihood("occu", X, W, y, device=rank, max_chains=64) as lk:
            sharded.attach_hybrid(lk, dist, rank, world, 2, 64, mode=mode)
            lp, gr = lk.logp_and_grad(th)
            err = sharded.comm_error(lk)
        ref = None
        if site_rank == 0:
            with bb.OccupancyLikelihood("occu", data["site_covs"], data["obs_covs"], data["obs"], device=rank) as full:
                ref = full.logp_and_grad(th)
        q.put((rank, g, lp, gr, err, ref))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_hybrid_chains_by_sites_grid(mode):
    """SURVEY 8e third row: 4 ranks = 2 chain groups x 2 site shards; the exchange stays inside a site group, each
    group's result equals the unsharded evaluation of ITS chains and is bit-identical on its two ranks."""
    if _n_gpus() < 4:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, 4, port, mode, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for g in (0, 1):
        a, b = res[2 * g], res[2 * g + 1]
        assert a[4] == 0 and b[4] == 0
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), "ranks of a site group disagree"
        np.testing.assert_allclose(a[2], a[5][0], rtol=2e-6)
        np.testing.assert_allclose(a[3], a[5][1], rtol=1e-5, atol=1e-5 * np.abs(a[5][1]).max())
    assert not np.array_equal(res[0][2], res[2][2]), "the two chain groups ran different chains"
