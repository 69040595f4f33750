"""Multi-GPU protocol of the hot path: shard by video, exchange only tiny BatchNorm statistics during the
step, all-reduce ONE flat gradient buffer at the end.

  reference (train.py:283-286)                   here
  -------------------------------------------    --------------------------------------------------------
  DistributedSampler: videos split over ranks    shard_videos(): contiguous videos per rank
  SyncBatchNorm fwd: all_gather(mean,invstd,n)   sync_stats_(): all_reduce(SUM) of [sum x, sum x^2] (float64),
                                                 n_global = local rows * world
  SyncBatchNorm bwd: all_reduce(sum dy, sum dy*xmu)  sync_stats_() on [sum dy, sum dy*xhat]; d(gamma), d(beta)
                                                 stay LOCAL sums (as torch does) and are averaged with the rest
  DDP bucketed all-reduce, mean over ranks       finish_flat_grads_(): one all_reduce(SUM) of the flat buffer,
                                                 scale 1/world applied while scattering to parameters
  per-rank loss = sum_local / sum_local_masks    unchanged (mean of per-rank means, NOT a global-mask mean)

The helpers work on any backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_videos(n_global: int, rank_: int, world: int) -> Tuple[int, int]:
    """[start, end) of the videos owned by `rank_`: equal contiguous shards (the global batch must divide)."""
    if n_global % world != 0:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank_ * per, (rank_ + 1) * per


def sync_stats_(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a BatchNorm statistics buffer (float64 [2*C])."""
    if world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def bn_global_rows(local_rows: int, world: int) -> int:
    return local_rows * max(world, 1)


def finish_flat_grads_(flat: torch.Tensor, group=None) -> float:
    """In-place SUM all-reduce of the flat gradient buffer; returns the scale (1/world) the caller applies
    when scattering into per-parameter gradients -- DDP's mean."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / w
    return 1.0
