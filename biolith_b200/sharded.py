"""Multi-GPU host plumbing (one process per GPU).

* chain sharding (default): every rank holds the whole dataset and its own chains -- no data-path
  collective; ``gather_chain_results`` concatenates the per-rank draws on rank 0 at the end.
* site sharding: ``shard_range`` gives each rank a contiguous block of sites; ``attach_site_sharding``
  wires the handle so that every evaluation sums the ranks' fp64 partial sums either with one
  ``ncclAllReduce`` (mode "nccl") or with the fused CUDA-IPC peer-memory kernel (mode "p2p").

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
