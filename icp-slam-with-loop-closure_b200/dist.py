"""Sharding of the pair list across ranks and the one exchange step of the path.

Scan pairs are independent: every rank holds the whole scan table, aligns
an interleaved-block slice of the problem index space, and an all-gather of fixed-size records
returns every pair's constraint to all ranks (rank 0 feeds the unchanged host-side pose graph).
This replaces the result gather of the reference's joblib fan-out
(``zip(*parallel(...))``, reference scripts/main.py:241).

One process per GPU; ``torch.distributed`` with the nccl backend on GPUs (gloo works for the
host-side logic and is what the CPU tests use).
"""
from __future__ import annotations

import numpy as np

RECORD_WIDTH = 8     # T (6 doubles), error, passes


def shard_indices(n_problems: int, rank: int, world: int, block: int = 4096) -> np.ndarray:
    """Global problem indices owned by `rank`: blocks rank, rank+world, ... of `block` problems.
    Interleaving evens out the data-dependent pass counts (4..102 per pair)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    block = max(int(block), 1)
    k = np.arange(n_problems, dtype=np.int64)
    return k[(k // block) % world == rank]


def shard_all_pairs(n_problems: int, rank: int, world: int, block: int = 4096):
    """(k_first, k_block, k_stride, B_local) for the device-decoded all-pairs index space
    (icpb_params pair_mode 1); enumerates exactly shard_indices(...)."""
    block = max(int(block), 1)
    b_local = int(len(shard_indices(n_problems, rank, world, block)))
    return rank * block, block, world * block, b_local


def pack_records(T: np.ndarray, error: np.ndarray, passes: np.ndarray) -> np.ndarray:
    rec = np.empty((len(error), RECORD_WIDTH))
    rec[:, :6] = np.asarray(T)[:, :2, :].reshape(-1, 6)
    rec[:, 6] = error
    rec[:, 7] = passes
    return rec


def unpack_records(rec: np.ndarray):
    T = np.zeros((len(rec), 3, 3))
    T[:, :2, :] = rec[:, :6].reshape(-1, 2, 3)
    T[:, 2, 2] = 1.0
    return T, rec[:, 6].copy(), rec[:, 7].astype(np.int32)


def all_gather_records(local: np.ndarray, n_problems: int, block: int = 4096, group=None, device=None):
    """All-gather per-rank record arrays (rows in shard_indices order) and return the
    (n_problems, 8) array in global problem order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_indices(n_problems, r, world, block)) for r in range(world)]
    if len(local) != counts[rank]:
        raise ValueError(f"rank {rank} holds {len(local)} records, expected {counts[rank]}")
    cap = max(counts) if counts else 0
    send = torch.zeros((cap, RECORD_WIDTH), dtype=torch.float64, device=device)
    if len(local):
        send[:len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(send.device)
    recv = torch.empty((world * cap, RECORD_WIDTH), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap, RECORD_WIDTH)
    out = np.empty((n_problems, RECORD_WIDTH))
    for r in range(world):
        out[shard_indices(n_problems, r, world, block)] = recv[r, :counts[r]]
    return out


def icp_batch_sharded(align, pairs, init_transforms=None, block: int = 4096, group=None, device=None,
                      **icp_kwargs):
    """Align this rank's slice with `align(pairs, init, **kw) -> BatchResult-like` (normally
    ``IcpEngine.run``) and return (T, error, passes) for ALL pairs on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    pairs = np.asarray(pairs)
    mine = shard_indices(len(pairs), rank, world, block)
    init = None if init_transforms is None else np.asarray(init_transforms)[mine]
    res = align(pairs[mine], init, **icp_kwargs)
    rec = all_gather_records(pack_records(res.T, res.error, res.iters), len(pairs), block, group, device)
    return unpack_records(rec)
