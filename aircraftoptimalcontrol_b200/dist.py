"""Multi-GPU plumbing: instances are independent, so they are split across ranks with NO collective on the hot
path; one all_gather at the end collects per-instance convergence and cost statistics.

One process per GPU (torchrun); `torch.distributed` backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_total: int, rank: int, world: int) -> np.ndarray:
    """Round-robin assignment (instance i -> rank i % world): parameter-sorted batches would otherwise give some
    ranks all the slow-converging instances.  Sizes differ by at most one."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def shard_sizes(n_total: int, world: int):
    return [(n_total - r + world - 1) // world for r in range(world)]


def pack_stats(stats: dict) -> np.ndarray:
    """(n,4) float64: iters, status, J, descent -- 32 B per instance."""
    return np.stack([stats["iters"].astype(np.float64), stats["status"].astype(np.float64),
                     np.asarray(stats["J"], dtype=np.float64), np.asarray(stats["descent"], dtype=np.float64)], axis=1)


def gather_stats(stats: dict, n_total: int, device=None):
    """All-gather the per-instance statistics of every rank's shard and put them back in global instance order.

    Returns dict(iters, status, J, descent) of length n_total on every rank.  With an uninitialised process
    group (single process) this is the identity."""
    import torch
    import torch.distributed as dist

    local = pack_stats(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allp, world = [local], 1
    else:
        world = dist.get_world_size()
        sizes = shard_sizes(n_total, world)
        nmax = max(sizes)
        dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
        buf = torch.zeros((nmax, 4), dtype=torch.float64, device=dev)
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
        out = torch.empty((world * nmax, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, buf)
        out = out.cpu().numpy().reshape(world, nmax, 4)
        allp = [out[r, : sizes[r]] for r in range(world)]
    full = np.zeros((n_total, 4))
    for r, part in enumerate(allp):
        full[shard_indices(n_total, r, world)] = part
    return dict(iters=full[:, 0].astype(np.int32), status=full[:, 1].astype(np.int32), J=full[:, 2], descent=full[:, 3])


def summarize(g: dict) -> dict:
    from ._lib import INST_CONVERGED
    return dict(total_iters=int(g["iters"].sum()), converged=int((g["status"] == INST_CONVERGED).sum()),
                max_iters=int(g["iters"].max()), mean_J=float(np.mean(g["J"])))
