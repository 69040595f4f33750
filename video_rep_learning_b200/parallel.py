"""Multi-GPU protocol of the hot path: shard by video, exchange only tiny BatchNorm statistics during the
step, all-reduce ONE flat gradient buffer at the end.

  reference (train.py:283-286)                   here
  -------------------------------------------    --------------------------------------------------------
  DistributedSampler: videos split over ranks    shard_videos(): contiguous videos per rank
  SyncBatchNorm fwd: all_gather(mean,invstd,n)   sync_stats_(): all_reduce(SUM) of [sum x, sum x^2] (float64),
                                                 n_global = local rows * world
  SyncBatchNorm bwd: all_reduce(sum dy, sum dy*xmu)  sync_stats_() on [sum dy, sum dy*xhat]; d(gamma), d(beta)
                                                 stay LOCAL sums (as torch does) and are averaged with the rest
  DDP bucketed all-reduce, mean over ranks       finish_flat_grads_(): one SUM of the flat buffer (in symmetric memory:
                                                 multimem / peer kernel of csrc/peer.cu; else all_reduce), scale 1/world
                                                 applied while scattering to parameters
  per-rank loss = sum_local / sum_local_masks    unchanged (mean of per-rank means, NOT a global-mask mean)

The helpers work on any backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
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
        self.max_n = lib.mvf_peer_buffer_bytes() // (2 * 16 * 16)   # 2 parities x 16 ranks x slot of 16-byte entries
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


class PeerFlatGrads:
    """The flat gradient buffer of a step in symmetric memory, summed over the ranks in place by csrc/peer.cu
    (mvf_peer_allreduce_f32: NVSwitch multimem.ld_reduce / multimem.st when the buffer has a multicast mapping, peer loads
    and stores otherwise) instead of an NCCL all-reduce.  One object per (group, device, size); `flat` is the SAME tensor
    every step (flat_grad_buffer zeroes it), which also makes the launch CUDA-graph replayable."""

    _cache = {}
    _failed = set()

    def __init__(self, group, device: torch.device, elems: int):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self._L = L
        lib = L.lib()
        pg = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        self.elems = (elems + 3) // 4 * 4
        self.flag_off = (self.elems * 4 + 255) // 256 * 256
        self.flag_bytes = lib.mvf_peer_allreduce_flag_bytes()
        nbytes = self.flag_off + 2 * self.flag_bytes
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, pg)
        self.ptrs_dev = int(self.handle.buffer_ptrs_dev)
        self.mc_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        # measured (scripts/peer_bench.py): with two ranks plain peer loads / stores from 64 CTAs beat the switch reduction
        # (48 vs 65 us for 19 MB; NCCL 58); from four ranks on the multicast path moves 1/world of the bytes per rank
        mode = os.environ.get("MVF_PEER_MULTICAST", "auto")
        if mode == "0" or (mode == "auto" and self.world <= 2):
            self.mc_ptr = 0
        self.ctas = int(os.environ.get("MVF_PEER_AR_CTAS", "0")) or (8 if self.mc_ptr else 64)   # multimem: more CTAs are slower (8: 71 us, 16: 82 us at 8 GPUs)
        self.counters = torch.zeros(2 * 64, dtype=torch.int32, device=device)     # two channels (see sum_range)
        self.flat = self.buf[:elems * 4].view(torch.float32)
        torch.cuda.synchronize(device)
        dist.barrier(group=pg)          # every rank's flags and counters are zero before the first exchange

    def owns(self, t: torch.Tensor) -> bool:
        return self.range_of(t) is not None

    def range_of(self, t: torch.Tensor):
        """(first, count) in floats when `t` is the buffer or a 16-byte aligned piece of it (a piece that reaches the end of
        the buffer takes the zero padding with it), else None."""
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.flat.device:
            return None
        off = t.data_ptr() - self.flat.data_ptr()
        if off < 0 or off % 16 != 0 or off // 4 + t.numel() > self.flat.numel():
            return None
        lo, n = off // 4, t.numel()
        if lo + n == self.flat.numel():
            n = self.elems - lo
        return (lo, n) if n % 4 == 0 else None

    def sum_range(self, lo: int, n: int, channel: int = 0, ctas: int = 0) -> None:
        """In-place SUM over the ranks of floats [lo, lo + n) on the current stream.  `channel` (0 / 1) selects the flag and
        counter set: two exchanges that may be in flight at the same time (the chain's gradients beside the pooling
        backward, the pooling gradients after it) must not share one."""
        L = self._L
        L.check(L.lib().mvf_peer_allreduce_f32(self.mc_ptr, self.ptrs_dev, 4 * lo, self.flag_off + channel * self.flag_bytes, n,
                                               self.rank, self.world, self.counters.data_ptr() + channel * 64 * 4,
                                               ctas or self.ctas, torch.cuda.current_stream(self.buf.device).cuda_stream),
                "mvf_peer_allreduce_f32")

    @classmethod
    def find(cls, t: torch.Tensor):
        """(object, first, count) of the symmetric buffer `t` is (a piece of), or None."""
        if t.is_cuda:
            for obj in cls._cache.values():
                r = obj.range_of(t)
                if r is not None:
                    return (obj,) + r
        return None

    @classmethod
    def get(cls, group, device: torch.device, elems: int):
        """The buffer object of (group, device, elems), created on first use (a collective: every rank reaches it in the
        same step); None when symmetric memory cannot be set up, during graph capture before the first eager step, or with
        MVF_PEER_AR=0 -> plain buffer + NCCL all-reduce."""
        key = (id(group) if group is not None else 0, device.index, elems)
        if key in cls._failed or os.environ.get("MVF_PEER_AR", "1") == "0":
            return None
        obj = cls._cache.get(key)
        if obj is None:
            if torch.cuda.is_current_stream_capturing():
                return None
            try:
                obj = cls(group, device, elems)
                cls._cache[key] = obj
            except Exception as e:          # pragma: no cover - depends on the machine
                import warnings
                warnings.warn(f"symmetric-memory gradient all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
                cls._failed.add(key)
                return None
        return obj


def flat_grad_buffer(elems: int, device: torch.device, group=None, enabled: bool = True) -> torch.Tensor:
    """Zeroed flat fp32 gradient buffer of a step: the symmetric-memory one when several ranks will all-reduce it on the
    NCCL backend, a fresh allocation otherwise."""
    if enabled and device.type == "cuda" and world_size(group) > 1 and dist.get_backend(group) == "nccl":
        obj = PeerFlatGrads.get(group, device, elems)
        if obj is not None:
            obj.flat.zero_()
            return obj.flat
    return torch.zeros(elems, dtype=torch.float32, device=device)


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


def finish_flat_grads_(flat: torch.Tensor, group=None, channel: int = 0, ctas: int = 0) -> float:
    """In-place SUM all-reduce of the flat gradient buffer (or of a piece of it); returns the scale (1/world) the caller
    applies when scattering into per-parameter gradients -- DDP's mean."""
    w = world_size(group)
    if w > 1:
        hit = PeerFlatGrads.find(flat)
        if hit is not None:
            hit[0].sum_range(hit[1], hit[2], channel, ctas)
            return 1.0 / w
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / w
    return 1.0
