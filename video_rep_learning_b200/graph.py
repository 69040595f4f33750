"""CUDA-graph replay of one MV-Former training step.

The eager step (models.TransformerModel.forward_tokens -> algos.SCL.compute_sequence_loss -> loss.backward()) enqueues
~125 kernels through six C-ABI calls plus PyTorch's autograd bookkeeping: ~1.8 ms of host time against ~2.2 ms of
device time at BASELINE configs[1], and every extra collective of the multi-GPU protocol (BatchNorm statistics, the flat
gradient all-reduce) adds host latency on top.  `GraphedTrainStep` captures exactly that launch sequence ONCE --
including the library's side-stream fork/join, the cross-rank exchanges (symmetric-memory kernels or NCCL) and the gradient scatter -- and replays it with one
`cudaGraphLaunch` per step.  Nothing about the math changes: the same kernels run in the same order on the same
buffers, which is what tests/test_gpu_graph.py checks (graph replay == eager step, bit for bit, dropout included).

What has to be static for a replay, and how it is kept so:
  * inputs live in buffers owned by this object (`tokens`, `seq_lens`, `steps`, `masks`); `__call__` copies new inputs
    in (or the producer writes into `tokens` directly -- zero copy);
  * parameters are read through raw pointers, so in-place optimizer updates are picked up; re-assigning a parameter
    tensor needs a new capture;
  * `.grad` of every head parameter is a view of one flat buffer in the graph's private pool, rewritten by each replay;
  * dropout: kernel arguments are frozen at capture, so the per-step seed is `base + *seed_dev` with `seed_dev` a device
    counter advanced inside the graph (csrc/common.cuh DropSeed) -- every replay draws a fresh mask.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib as L
from . import engine


class GraphedTrainStep:
    def __init__(self, model, algo, Bv: int, T: int, P: int, C_in: int, dtype=torch.bfloat16,
                 device: Optional[torch.device] = None, warmup: int = 3, project: bool = True, optimizer=None):
        self.model, self.algo = model, algo
        self.Bv, self.T = Bv, T
        self.project = project
        dev = device if device is not None else next(model.embed.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs a CUDA device: the MV-Former hot path has no CPU implementation")
        self.device = dev
        self.tokens = torch.zeros(2 * Bv, T, P, C_in, dtype=dtype, device=dev)
        self.seq_lens = torch.ones(Bv, 2, dtype=torch.int64, device=dev)
        self.steps = torch.zeros(Bv, 2, T, dtype=torch.int64, device=dev)
        self.masks = torch.ones(2 * Bv, 1, T, dtype=torch.float32, device=dev)
        self.params: List[torch.nn.Parameter] = [p for n, p in model.named_parameters() if "backbone" not in n]
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.warmup = warmup
        # optional fused optimizer tail (optim.FusedAdam): clip + Adam captured at the end of the same graph
        self.optimizer = optimizer
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self.grads: List[Optional[torch.Tensor]] = []
        self._scratch_refs: List[torch.Tensor] = []
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------------
    def set_inputs(self, tokens=None, seq_lens=None, steps=None, masks=None):
        """Copy a step's inputs into the static buffers (async on the current stream; host tensors should be pinned)."""
        if tokens is not None and tokens.data_ptr() != self.tokens.data_ptr():
            self.tokens.copy_(tokens.reshape(self.tokens.shape), non_blocking=True)
        if seq_lens is not None:
            self.seq_lens.copy_(seq_lens.reshape(self.seq_lens.shape), non_blocking=True)
        if steps is not None:
            self.steps.copy_(steps.reshape(self.steps.shape), non_blocking=True)
        if masks is not None:
            self.masks.copy_(masks.reshape(self.masks.shape), non_blocking=True)

    def adopt_tokens(self, tokens: torch.Tensor):
        """Use `tokens` itself as the static token buffer (before capture): the producer writes there, nothing is copied."""
        if self.graph is not None:
            raise RuntimeError("adopt_tokens() must be called before capture()")
        if tuple(tokens.shape) != tuple(self.tokens.shape) or tokens.dtype != self.tokens.dtype or not tokens.is_contiguous():
            raise ValueError("token buffer must be contiguous with the shape / dtype this step was built for")
        self.tokens = tokens

    def _eager_step(self) -> torch.Tensor:
        embs = self.model.forward_tokens(self.tokens, video_masks=self.masks, project=self.project)
        loss = self.algo.compute_sequence_loss(embs.view(self.Bv, 2, self.T, -1), self.seq_lens, self.steps, self.masks)["loss"]
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def capture(self, profile: int = 0) -> "GraphedTrainStep":
        """Warm up eagerly on a side stream (lazy one-time work inside the library: kernel attributes, side streams,
        scratch buffers, NCCL communicators), then capture one step.  profile=True keeps the library's CUDA-event
        brackets around the dominant kernels in the graph (external event-record nodes, read with mvf_profile_read
        after a replay)."""
        embed = self.model.embed
        embed.seed_dev = self.seed_dev
        return self._capture(profile)

    def _snapshot_state(self):
        """Everything the eager warm-up steps mutate besides scratch: parameters (optimizer), BatchNorm running statistics
        and counters, the optimizer's moments and step counter.  Restored after the warm-up, so capturing is free of side
        effects on the training state whatever is in the static input buffers at that point."""
        snap = dict(params=[p.detach().clone() for p in self.params],
                    buffers=[(b, b.detach().clone()) for n, b in self.model.named_buffers() if "backbone" not in n])
        if self.optimizer is not None:
            snap["opt"] = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.optimizer.state.get(p, {}).items()}
                           for p in self.params}
            snap["opt_step"] = getattr(self.optimizer, "step_count", None)
        return snap

    def _restore_state(self, snap):
        with torch.no_grad():
            for p, v in zip(self.params, snap["params"]):
                p.copy_(v)
            for b, v in snap["buffers"]:
                b.copy_(v)
            if self.optimizer is not None:
                for p in self.params:
                    st, old = self.optimizer.state.get(p), snap["opt"].get(id(p), {})
                    if not st:
                        continue
                    for k, v in st.items():
                        if torch.is_tensor(v):               # in place: the captured launches hold these pointers
                            v.copy_(old[k]) if k in old else v.zero_()
                if snap.get("opt_step") is not None and hasattr(self.optimizer, "set_step_count"):
                    self.optimizer.set_step_count(snap["opt_step"])

    def _capture(self, profile: bool) -> "GraphedTrainStep":
        cur = torch.cuda.current_stream(self.device)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(cur)
        snap = self._snapshot_state()
        with torch.cuda.stream(s):
            for _ in range(max(self.warmup, 1)):
                for p in self.params:
                    p.grad = None
                self._eager_step()
            self._restore_state(snap)
        cur.wait_stream(s)
        torch.cuda.synchronize(self.device)
        del snap
        for p in self.params:
            p.grad = None
        lib = L.lib()
        if profile:
            lib.mvf_profile_enable(int(profile))   # 1: pooling kernels only, 2: every tagged kernel group (mvf_profile_enable)
        n0 = lib.mvf_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.seed_dev.add_(1)
            self.loss = self._eager_step()
        self.launches_per_step = int(lib.mvf_launch_count() - n0)
        self.graph = g
        self.grads = [p.grad for p in self.params]
        # the captured launches hold raw pointers into the engine's cached scratch workspaces: keep those tensors alive for as
        # long as the graph is, even if a later eager call (e.g. whole-video evaluation) makes the cache grow a replacement
        self._scratch_refs = engine.scratch_tensors(self.device)
        return self

    def __call__(self, tokens=None, seq_lens=None, steps=None, masks=None) -> torch.Tensor:
        """Run one step.  Returns the (static) loss tensor; `.grad` of the head parameters holds this step's gradients."""
        self.set_inputs(tokens, seq_lens, steps, masks)
        if self.graph is None:
            self.capture()
        if self.optimizer is not None and hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()       # a scheduler may have changed param_groups[...]['lr'] since the last replay
        self.graph.replay()
        for p, gr in zip(self.params, self.grads):
            if p.grad is not gr:
                p.grad = gr            # an optimizer's zero_grad(set_to_none=True) detaches them; re-attach the static views
        return self.loss

    def release(self):
        self.graph = None
        self.loss = None
        self.grads = []
        self._scratch_refs = []
        if self.model.embed.seed_dev is self.seed_dev:
            self.model.embed.seed_dev = None
