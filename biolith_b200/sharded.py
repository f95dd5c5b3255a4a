"""Multi-GPU host plumbing (one process per GPU).

* chain sharding (default): every rank holds the whole dataset and its own chains -- no data-path
  collective; ``gather_chain_results`` concatenates the per-rank draws on rank 0 at the end.
* site sharding: ``shard_range`` gives each rank a contiguous block of sites; ``attach_site_sharding``
  wires the handle so that every evaluation sums the ranks' fp64 partial sums either with one
  ``ncclAllReduce`` (mode "nccl") or with the fused CUDA-IPC peer-memory kernel (mode "p2p").
* hybrid chains x sites grid (third row): ``hybrid_layout`` cuts the world into site groups of
  ``site_group_size`` consecutive ranks; the ranks of a group hold the site shards of ONE dataset and share the
  same chains (exchange inside the group only), different groups run different chains and never talk.
  ``attach_hybrid`` builds the per-group sub-communicator and attaches it.

``dist`` is an initialised ``torch.distributed`` module (or any object with the same
``broadcast_object_list`` / ``all_gather_object`` / ``gather_object`` functions): torch is only the
rendezvous here, it is never on the data path and this package never imports it.
"""

from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from ._lib import check

NCCL_UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


def shard_range(n_sites: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced site block of ``rank`` (any split is valid: the log-density is a sum over sites)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_sites), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_data(site_covs, obs_covs, obs, session_duration, rank: int, world: int):
    """Slice reference-layout arrays to this rank's sites."""
    lo, hi = shard_range(np.asarray(site_covs).shape[0], rank, world)
    sd = None if session_duration is None else np.asarray(session_duration)[lo:hi]
    obs = np.asarray(obs)
    obs = obs[:, lo:hi] if obs.ndim == 4 else obs[lo:hi]
    return np.asarray(site_covs)[lo:hi], np.asarray(obs_covs)[lo:hi], obs, sd


def exchange_unique_id(dist, rank: int, make_id) -> bytes:
    """rank 0 creates the 128-byte NCCL id, everyone receives it."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) != NCCL_UNIQUE_ID_BYTES:
        raise RuntimeError("NCCL unique id exchange failed")
    return bytes(box[0])


def attach_site_sharding(lk, dist, rank: int, world: int, max_chains: int, mode: str = "nccl"):
    """After this call ``lk`` evaluates the log-density of ALL ranks' site shards (identical on every rank)."""
    lib = _lib.load()
    if mode == "nccl":
        def make_id():
            buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
            check(lib.bl_comm_unique_id(buf, NCCL_UNIQUE_ID_BYTES), "bl_comm_unique_id")
            return buf.raw

        uid = exchange_unique_id(dist, rank, make_id)
        check(lib.bl_dataset_attach_nccl(lk.handle, uid, len(uid), rank, world), "bl_dataset_attach_nccl")
    elif mode == "p2p":
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        check(lib.bl_dataset_p2p_export(lk.handle, rank, world, max_chains, buf, IPC_HANDLE_BYTES),
              "bl_dataset_p2p_export")
        handles = [None] * world
        dist.all_gather_object(handles, buf.raw)
        blob = b"".join(handles)
        check(lib.bl_dataset_p2p_attach(lk.handle, blob, IPC_HANDLE_BYTES), "bl_dataset_p2p_attach")
        dist.barrier()
    else:
        raise ValueError("mode must be 'nccl' or 'p2p'")
    return mode


class _SubGroup:
    """The slice of a ``torch.distributed``-like module that ``attach_site_sharding`` needs, restricted to one
    process group whose members are the global ranks ``members`` (group rank = position in that list)."""

    def __init__(self, dist, group, members):
        self._dist, self._group, self._members = dist, group, list(members)

    def broadcast_object_list(self, box, src=0):
        self._dist.broadcast_object_list(box, src=self._members[src], group=self._group)

    def all_gather_object(self, out, obj):
        self._dist.all_gather_object(out, obj, group=self._group)

    def barrier(self):
        self._dist.barrier(group=self._group)


def hybrid_layout(rank: int, world: int, site_group_size: int):
    """-> (chain_group, site_rank, members): ``world`` = n_chain_groups x site_group_size; ranks
    [g * size, (g + 1) * size) form site group g (consecutive ranks = neighbouring GPUs)."""
    if site_group_size < 1 or world % site_group_size != 0:
        raise ValueError("world must be a multiple of site_group_size")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    g = rank // site_group_size
    return g, rank % site_group_size, list(range(g * site_group_size, (g + 1) * site_group_size))


def attach_hybrid(lk, dist, rank: int, world: int, site_group_size: int, max_chains: int, mode: str = "nccl"):
    """Chains x sites grid: after this call ``lk`` (holding site shard ``site_rank`` of its group's dataset)
    evaluates the log-density of the whole group's sites; groups are independent (their own chains, no traffic
    between them).  Every rank must call it (``new_group`` is collective over the world).  Returns
    ``(chain_group, site_rank, subgroup)``."""
    groups = []
    n_groups = world // site_group_size
    for g in range(n_groups):  # every rank creates every group, in the same order (torch.distributed contract)
        members = list(range(g * site_group_size, (g + 1) * site_group_size))
        groups.append((dist.new_group(ranks=members), members))
    g, site_rank, members = hybrid_layout(rank, world, site_group_size)
    sub = _SubGroup(dist, groups[g][0], members)
    if site_group_size > 1:
        attach_site_sharding(lk, sub, site_rank, site_group_size, max_chains, mode=mode)
    return g, site_rank, sub


def comm_error(lk) -> int:
    err = C.c_int32(0)
    check(_lib.load().bl_dataset_comm_error(lk.handle, C.byref(err)), "bl_dataset_comm_error")
    return int(err.value)


def gather_chain_results(dist, rank: int, world: int, local: dict):
    """Chain sharding: concatenate per-rank NUTS results along the chain axis on rank 0."""
    boxes = [None] * world if rank == 0 else None
    dist.gather_object(local, boxes, dst=0)
    if rank != 0:
        return None
    out = {}
    for k, v in boxes[0].items():
        if isinstance(v, np.ndarray) and v.ndim >= 1:
            out[k] = np.concatenate([b[k] for b in boxes], axis=0)
        else:
            out[k] = [b[k] for b in boxes]
    return out
