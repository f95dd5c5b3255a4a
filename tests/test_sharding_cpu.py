"""CPU (gloo, world_size 2) tests of the multi-GPU host plumbing: site ranges, the NCCL-id /
IPC-handle exchange pattern and the chain gather.  No GPU, no compute calls."""

import os
import socket
import sys

import numpy as np
import pytest

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


def test_shard_ranges_partition_the_sites():
    from biolith_b200.sharded import shard_data, shard_range

    for n in (0, 1, 7, 8, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    X = np.arange(10)[:, None].astype(float)
    W = np.zeros((10, 1, 3, 2))
    y = np.zeros((1, 10, 1, 3))
    a = shard_data(X, W, y, None, 0, 3)
    assert a[0].shape[0] == 4 and a[1].shape[0] == 4 and a[2].shape == (1, 4, 1, 3)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from biolith_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = sharded.exchange_unique_id(dist, rank, lambda: bytes(range(128)))
        handles = [None] * world
        dist.all_gather_object(handles, bytes([rank]) * sharded.IPC_HANDLE_BYTES)
        local = dict(samples=np.full((3, 5, 2), float(rank)), step_size=np.full(3, 0.1 * (rank + 1)), wall_s=1.0)
        g = sharded.gather_chain_results(dist, rank, world, local)
        q.put((rank, uid == bytes(range(128)), [h[0] for h in handles], None if g is None else g["samples"].shape,
               None if g is None else g["samples"][:, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_and_gather():
    import torch.multiprocessing as mp

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == [0, 1] and res[1][2] == [0, 1]
    assert res[0][3] == (6, 5, 2) and res[0][4] == [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    assert res[1][3] is None


def test_hybrid_layout_is_a_grid():
    from biolith_b200.sharded import hybrid_layout

    seen = {}
    for r in range(8):
        g, sr, members = hybrid_layout(r, 8, 2)
        assert members == [2 * g, 2 * g + 1] and members[sr] == r
        seen.setdefault(g, []).append(sr)
    assert seen == {0: [0, 1], 1: [0, 1], 2: [0, 1], 3: [0, 1]}
    assert hybrid_layout(5, 8, 8)[:2] == (0, 5) and hybrid_layout(5, 8, 1)[:2] == (5, 0)
    with pytest.raises(ValueError):
        hybrid_layout(0, 6, 4)


def _hybrid_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from biolith_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the sub-communicator plumbing of attach_hybrid without a GPU handle: the NCCL-id / IPC-handle exchange
        # pattern runs inside each site group only
        groups = [(dist.new_group(ranks=[2 * g, 2 * g + 1]), [2 * g, 2 * g + 1]) for g in range(world // 2)]
        g, site_rank, members = sharded.hybrid_layout(rank, world, 2)
        sub = sharded._SubGroup(dist, groups[g][0], members)
        uid = sharded.exchange_unique_id(sub, site_rank, lambda: bytes([g]) * 128)
        handles = [None] * 2
        sub.all_gather_object(handles, bytes([rank]) * sharded.IPC_HANDLE_BYTES)
        sub.barrier()
        q.put((rank, g, site_rank, uid[0], [h[0] for h in handles]))
    finally:
        dist.destroy_process_group()


def test_gloo_world4_hybrid_subgroups():
    import torch.multiprocessing as mp

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank got ITS group's id (made by the group's rank 0) and only its group's handles
    assert [(r[1], r[2], r[3], r[4]) for r in res] == [(0, 0, 0, [0, 1]), (0, 1, 0, [0, 1]), (1, 0, 1, [2, 3]),
                                                       (1, 1, 1, [2, 3])]
