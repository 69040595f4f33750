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


class PeerStats:
    """Cross-rank sum of small float64 buffers over NVLink peer memory (csrc/peer.cu) instead of an NCCL all-reduce:
    one symmetric buffer per rank (torch.distributed._symmetric_memory allocates it and exchanges the mappings), one
    small kernel per rank and exchange, graph-replayable (the exchange counter lives on the device)."""

    _cache = {}
    _failed = set()

    def __init__(self, group, device: torch.device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self._L, self._C = L, C
        lib = L.lib()
        pg = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        self.max_n = (lib.mvf_peer_buffer_bytes() - 256) // 16
        self.buf = symm_mem.empty(lib.mvf_peer_buffer_bytes(), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, pg)
        self.ptrs_dev = int(self.handle.buffer_ptrs_dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=pg)          # every rank's flags and counter are zero before the first exchange

    def sum_(self, stats: torch.Tensor) -> torch.Tensor:
        L = self._L
        L.check(L.lib().mvf_peer_sum_f64(stats.data_ptr(), stats.numel(), self.ptrs_dev, self.rank, self.world,
                                         self.counter.data_ptr(), torch.cuda.current_stream(stats.device).cuda_stream),
                "mvf_peer_sum_f64")
        return stats

    @classmethod
    def get(cls, group, device: torch.device):
        """The exchange object of (group, device), created on first use; None when it cannot be set up (no symmetric
        memory on this system, capture in progress before the first eager step, MVF_PEER_BN=0) -> NCCL all-reduce."""
        import os
        key = (id(group) if group is not None else 0, device.index)
        if key in cls._failed or os.environ.get("MVF_PEER_BN", "1") == "0":
            return None
        obj = cls._cache.get(key)
        if obj is None:
            if torch.cuda.is_current_stream_capturing():
                return None
            try:
                obj = cls(group, device)
                cls._cache[key] = obj
            except Exception as e:          # pragma: no cover - depends on the machine
                import warnings
                warnings.warn(f"NVLink peer exchange of BatchNorm statistics unavailable ({type(e).__name__}: {e}); "
                              "using NCCL all-reduce")
                cls._failed.add(key)
                return None
        return obj


def sync_stats_(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of a BatchNorm statistics buffer (float64 [2*C]) over the ranks: NVLink peer exchange for CUDA
    buffers on the NCCL backend (PeerStats), all-reduce otherwise (gloo in the CPU tests)."""
    if world_size(group) > 1:
        if stats.is_cuda and dist.get_backend(group) == "nccl":
            peer = PeerStats.get(group, stats.device)
            if peer is not None and stats.numel() <= peer.max_n and stats.dtype == torch.float64 and stats.is_contiguous():
                return peer.sum_(stats)
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
