"""Fused optimizer tail of the training step.

`train.py:124-133,151-155` ends every iteration with GradScaler.unscale_ (AMP), `clip_grad_norm_(GRAD_CLIP)` and
`optimizer.step()` of the Adam / AdamW built by `utils/optimizer.py:60-73` -- in PyTorch ~4 launches per parameter tensor
(56 tensors) plus a host sync for the clip factor.  `FusedAdam` does the same arithmetic in two launches over all tensors
(csrc/optim.cu), with the step count and the learning rate in device memory so that the tail can be captured at the end
of the step graph (`graph.GraphedTrainStep(..., optimizer=opt)`).  It is a `torch.optim.Optimizer`: param groups and `lr`
schedulers (`utils/optimizer.py:construct_scheduler`) work as usual -- under graph replay the new learning rate is copied to the
device scalar by `sync_lr()` before each replay -- and `state_dict()` / `load_state_dict()` carry the step count as a
per-parameter `step` entry, interchangeable with `torch.optim.Adam` checkpoints.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: float = 0.0, adamw: bool = False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        self.adamw = bool(adamw)
        self.grad_norm: Optional[torch.Tensor] = None      # pre-clip global gradient norm of the last step (device scalar)
        self._dev_state = {}
        self._pending_step: Optional[int] = None            # step count loaded from a checkpoint before the device state exists

    def _group_state(self, gi: int, device: torch.device):
        st = self._dev_state.get(gi)
        if st is None:
            st = dict(lr=torch.zeros(1, dtype=torch.float32, device=device), lr_host=None,
                      step=torch.zeros(1, dtype=torch.int64, device=device), ws=None)
            if self._pending_step is not None:
                st["step"].fill_(int(self._pending_step))
                self._pending_step = None
            self._dev_state[gi] = st
        return st

    def sync_lr(self):
        """Copy param_groups[0]['lr'] into the device-resident scalar the kernels read.  step() does this itself; a CUDA-graph
        replay never runs step(), so graph.GraphedTrainStep calls it before every replay (outside the captured region)."""
        st = self._dev_state.get(0)
        lr = self.param_groups[0]["lr"]
        if st is not None and st["lr_host"] != lr:
            st["lr"].fill_(float(lr))
            st["lr_host"] = lr

    # ---- checkpointing: the step count lives on the device; mirror it into the per-parameter state like torch.optim.Adam ----
    def set_step_count(self, t: int):
        st = self._dev_state.get(0)
        if st is None:
            self._pending_step = int(t)
        else:
            st["step"].fill_(int(t))

    def state_dict(self):
        t = float(self.step_count)
        for p, st in self.state.items():
            if "exp_avg" in st:
                st["step"] = torch.tensor(t, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # our own checkpoints and torch.optim.Adam / AdamW checkpoints both carry a per-parameter 'step'
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        if steps:
            self.set_step_count(int(round(max(steps))))

    @torch.no_grad()
    def step(self, closure=None, inv_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.lib()
        # the clip factor is global over every group (clip_grad_norm_(model.parameters())): one call per group would clip per
        # group, so all groups must share their hyper-parameters (utils/optimizer.py builds them that way) and go in one call
        tensors, group0 = [], self.param_groups[0]
        for g in self.param_groups:
            if (g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]) != (group0["lr"], tuple(group0["betas"]), group0["eps"],
                                                                              group0["weight_decay"]):
                raise NotImplementedError("FusedAdam: parameter groups with different hyper-parameters")
            tensors += [p for p in g["params"] if p.grad is not None]
        if not tensors:
            return loss
        dev = tensors[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA tensors only (there is no CPU implementation)")
        for p in tensors:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise TypeError("FusedAdam: parameters and gradients must be contiguous float32 tensors")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        gs = self._group_state(0, dev)
        if gs["lr_host"] != group0["lr"]:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: change the learning rate outside the captured region")
            gs["lr"].fill_(float(group0["lr"]))
            gs["lr_host"] = group0["lr"]
        n = len(tensors)
        need = lib.mvf_opt_ws_bytes(n)
        if gs["ws"] is None or gs["ws"].numel() < need:
            gs["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
        if self.grad_norm is None:
            self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        numel = (C.c_int64 * n)(*[p.numel() for p in tensors])
        b1, b2 = group0["betas"]
        with torch.cuda.device(dev):
            L.check(lib.mvf_opt_adam_step(n, arr(tensors), arr([p.grad for p in tensors]),
                                          arr([self.state[p]["exp_avg"] for p in tensors]),
                                          arr([self.state[p]["exp_avg_sq"] for p in tensors]), numel, gs["lr"].data_ptr(),
                                          gs["step"].data_ptr(), float(b1), float(b2), float(group0["eps"]),
                                          float(group0["weight_decay"]), 1 if self.adamw else 0, self.max_grad_norm,
                                          float(inv_scale), self.grad_norm.data_ptr(), gs["ws"].data_ptr(), gs["ws"].numel(),
                                          torch.cuda.current_stream(dev).cuda_stream), "mvf_opt_adam_step")
        return loss

    @property
    def step_count(self) -> int:
        st = self._dev_state.get(0)
        if st is None:
            return int(self._pending_step or 0)
        return int(st["step"].item())


def construct_optimizer(model, cfg) -> FusedAdam:
    """utils/optimizer.py:10-78 for the frozen-backbone MV-Former configs: every non-backbone parameter, one weight decay,
    Adam or AdamW, with train.py's clip_grad_norm_(OPTIMIZER.GRAD_CLIP) folded into the step."""
    params = [p for n, p in model.named_parameters() if "backbone" not in n and p.requires_grad]
    kind = cfg.OPTIMIZER.TYPE
    if kind not in ("AdamOptimizer", "AdamWOptimizer"):
        raise NotImplementedError(f"FusedAdam covers AdamOptimizer / AdamWOptimizer, not {kind}")
    return FusedAdam(params, lr=cfg.OPTIMIZER.LR.INITIAL_LR, betas=(0.9, 0.999), weight_decay=cfg.OPTIMIZER.WEIGHT_DECAY,
                     max_grad_norm=float(cfg.OPTIMIZER.GRAD_CLIP) if "GRAD_CLIP" in cfg.OPTIMIZER else 0.0,
                     adamw=(kind == "AdamWOptimizer"))
