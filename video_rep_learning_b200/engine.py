"""Host-side driver of the fused MV-Former head / projection / SCL kernels.

PyTorch is used here for what it is good at -- owning device memory, streams, autograd bookkeeping and
torch.distributed -- while every FLOP of the hot path runs inside libmvf_b200.so (see include/mvf_b200.h).

Three autograd Functions wrap the C ABI:
  * HeadFn   : tokens -> frame embeddings          (MultiEntityTransformerEmbModel.forward, mvformer.py:128-200)
  * ProjFn   : embeddings -> unit-norm projections (MLPHead + F.normalize, resnet_c2d.py:112-126, transformer.py:226-230)
  * ModelFn  : both of the above in one node, one flat gradient buffer, ONE all-reduce (replaces DDP's bucketed
               all-reduce, train.py:285-286) and optional cross-rank BatchNorm statistics (replaces SyncBatchNorm,
               train.py:283)
  * SCLFn    : SCL.compute_sequence_loss forward+backward in one launch sequence (algos/scl.py:52-105)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from . import parallel


# --------------------------------------------------------------------------------------------------
# static description of a head
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class HeadSpec:
    c_in: int
    n_entities: int = 3
    pool_channels: int = 384
    fc_channels: Tuple[int, ...] = (512, 512)
    hidden: int = 256
    d_ff: int = 1024
    n_heads: int = 8
    n_layers: int = 3
    emb: int = 128
    proj: int = 128
    one_hot: str = "pool"
    final: str = "one"
    train_frames: int = 20
    drop_p: float = 0.1
    ln_eps: float = 1e-5
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    pool_kind: str = "lstp"     # "lstp" LearnableTokenPooling | "fwb" FIXED_WIDTH_BASELINE (FWBPooling on the CLS embedding)
    cls_dim: int = 0            # fwb: width of the CLS embedding


@dataclass
class RunOptions:
    """Per-model switches that are not part of the reference cfg."""
    gemm_backend: int = L.GEMM_AUTO
    sync_bn: bool = True            # exchange BatchNorm statistics across ranks (reference: SyncBatchNorm)
    allreduce_grads: bool = True    # all-reduce the flat gradient buffer inside backward (reference: DDP)
    process_group: Optional[object] = None
    scl_quirk: bool = True          # keep the 1e-6 weight on masked columns (algos/scl.py:80)
    pool_mode: int = L.POOL_AUTO    # entity pooling: AUTO -> folded (no K|V tensors); POOL_DENSE = as written in the reference
    overlap_grad_allreduce: bool = True   # several ranks, symmetric-memory gradient buffer: the chain's gradients are summed
                                          # by a few CTAs (csrc/peer.cu) beside the pooling backward, the last ~0.25 ms of the
                                          # step; only the pooling gradients wait for its end
    pool_bwd_reserve_sms: int = 4         # SMs the pooling backward leaves to those CTAs (its persistent CTAs fill an SM)


def _world(opts: RunOptions) -> int:
    """Ranks of the process group: gates (and scales) the gradient all-reduce."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(opts.process_group)
    return 1


def _bn_world(opts: RunOptions) -> int:
    """Ranks whose rows enter one BatchNorm statistic: the descriptor's world_size (it divides the summed statistics by
    local_rows * world_size).  Without sync_bn every rank normalises with its own batch, whatever the group size."""
    return _world(opts) if opts.sync_bn else 1


class Plan:
    """Descriptor + buffer sizes + canonical parameter order for one (spec, shape, dtype, mode)."""

    _cache: Dict[tuple, "Plan"] = {}

    def __init__(self, spec: HeadSpec, BV: int, T: int, P: int, dtype: int, training: bool, has_mask: bool,
                 world: int, backend: int, pool_mode: int = L.POOL_AUTO):
        lib = L.lib()
        d = L.HeadDesc()
        d.BV, d.T, d.P, d.C_in = BV, T, P, spec.c_in
        d.E, d.SPC = spec.n_entities, spec.pool_channels
        if len(spec.fc_channels) > L.MAX_FC:
            raise NotImplementedError(f"FC_LAYERS with {len(spec.fc_channels)} layers (max {L.MAX_FC})")
        d.n_fc = len(spec.fc_channels)
        for i, ch in enumerate(spec.fc_channels):
            d.fc[i] = ch
        d.H, d.DFF, d.heads, d.L = spec.hidden, spec.d_ff, spec.n_heads, spec.n_layers
        d.D, d.PS = spec.emb, spec.proj
        d.one_hot = L.ONEHOT[spec.one_hot]
        d.final_mode = L.FINAL[spec.final]
        d.train_frames = spec.train_frames
        d.dtype = dtype
        d.training = 1 if training else 0
        d.has_mask = 1 if has_mask else 0
        d.gemm_backend = backend
        d.world_size = world
        d.pool_mode = pool_mode
        d.pool_kind = L.POOLKIND[spec.pool_kind]
        d.cls_dim = spec.cls_dim
        d.drop_p = spec.drop_p
        d.ln_eps, d.bn_eps, d.bn_momentum = spec.ln_eps, spec.bn_eps, spec.bn_momentum
        d.seed = 0
        self.desc = d
        self.spec = spec
        n = lib.mvf_num_params(C.byref(d))
        if n < 0:
            raise RuntimeError(f"mvf_num_params failed: {L.last_error()}")
        self.param_names: List[str] = []
        self.param_shapes: List[Tuple[int, int]] = []
        buf = C.create_string_buffer(256)
        r, c = C.c_int64(), C.c_int64()
        for i in range(n):
            L.check(lib.mvf_param_info(C.byref(d), i, buf, 256, C.byref(r), C.byref(c)), "mvf_param_info")
            self.param_names.append(buf.value.decode())
            self.param_shapes.append((r.value, c.value))
        self.n_head_params = sum(1 for nm in self.param_names if nm.startswith("embed."))
        self.bn_names: List[str] = []
        for i in range(lib.mvf_num_bn(C.byref(d))):
            L.check(lib.mvf_bn_info(C.byref(d), i, buf, 256, C.byref(r)), "mvf_bn_info")
            self.bn_names.append(buf.value.decode())
        self.save_bytes = lib.mvf_save_bytes(C.byref(d))
        self.ws_bytes = lib.mvf_ws_bytes(C.byref(d))
        self.gpack_elems = lib.mvf_gpack_elems(C.byref(d))
        self.gpack_pool_elems = lib.mvf_gpack_pool_elems(C.byref(d))
        self.proj_save_bytes = lib.mvf_proj_save_bytes(C.byref(d))
        self.proj_ws_bytes = lib.mvf_proj_ws_bytes(C.byref(d))
        self.n_fc = d.n_fc
        self.world = world

    @classmethod
    def get(cls, spec: HeadSpec, BV: int, T: int, P: int, dtype: int, training: bool, has_mask: bool, world: int,
            backend: int, pool_mode: int = L.POOL_AUTO) -> "Plan":
        key = (spec, BV, T, P, dtype, training, has_mask, world, backend, pool_mode)
        p = cls._cache.get(key)
        if p is None:
            p = cls(spec, BV, T, P, dtype, training, has_mask, world, backend, pool_mode)
            cls._cache[key] = p
        return p

    def desc_with_seed(self, seed: int, seed_dev: Optional[torch.Tensor] = None,
                       cls_emb: Optional[torch.Tensor] = None) -> L.HeadDesc:
        d = L.HeadDesc()
        C.memmove(C.byref(d), C.byref(self.desc), C.sizeof(L.HeadDesc))
        d.seed = seed
        d.seed_dev = None if seed_dev is None else seed_dev.data_ptr()
        d.cls_emb = None if cls_emb is None else cls_emb.data_ptr()
        return d

    def lookup(self, name: str):
        off, r, c, ld, dt = C.c_size_t(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        L.check(L.lib().mvf_save_lookup(C.byref(self.desc), name.encode(), C.byref(off), C.byref(r), C.byref(c),
                                        C.byref(ld), C.byref(dt)), "mvf_save_lookup")
        return off.value, r.value, c.value, ld.value, dt.value

    def region(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        """Typed [rows, cols] view of a named region of a save / gpack buffer (tests, attention maps)."""
        off, r, c, ld, dt = self.lookup(name)
        tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.float64, 3: torch.int32}[dt]
        esz = {0: 4, 1: 2, 2: 8, 3: 4}[dt]
        raw = buf.view(torch.uint8)[off: off + r * ld * esz].view(tdt).view(r, ld)
        return raw[:, :c]

    def bn_stat(self, buf: torch.Tensor, bn_idx: int, backward: bool) -> torch.Tensor:
        off, n = C.c_size_t(), C.c_int64()
        L.check(L.lib().mvf_bn_stat_lookup(C.byref(self.desc), bn_idx, 1 if backward else 0, C.byref(off), C.byref(n)),
                "mvf_bn_stat_lookup")
        return buf.view(torch.uint8)[off.value: off.value + 8 * n.value].view(torch.float64)


_ws_cache: Dict[tuple, torch.Tensor] = {}


def _scratch(device: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    key = (device, tag)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def scratch_tensors(device: torch.device) -> List[torch.Tensor]:
    """The cached scratch workspaces currently in use on `device` (graph.GraphedTrainStep pins them while its graph lives)."""
    return [t for (dev, _tag), t in _ws_cache.items() if dev == device]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _mvf_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.MVF_F32
    if t.dtype == torch.bfloat16:
        return L.MVF_BF16
    raise TypeError(f"tokens must be float32 or bfloat16, got {t.dtype}")


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: the MV-Former hot path has no CPU implementation")


@dataclass
class CallState:
    """Everything a Function needs besides tensors."""
    spec: HeadSpec
    opts: RunOptions
    training: bool
    bn_running: List[Optional[torch.Tensor]]      # [rm0, rv0, rm1, rv1, ..., rm_proj, rv_proj]
    bn_tracked: List[Optional[torch.Tensor]]
    project: int = 1                 # 1 MLPHead + normalise, 0 normalise only, 2 MLPHead only
    want_attn: bool = False
    seed: int = 0
    seed_dev: Optional[torch.Tensor] = None   # device-resident int64 counter added to `seed` in the kernels (graph replay)
    cls_emb: Optional[torch.Tensor] = None    # FIXED_WIDTH_BASELINE: [BV*T, cls_dim] fp32 CLS embeddings of the frames
    # filled by forward
    plan: Optional[Plan] = None
    head_save: Optional[torch.Tensor] = None
    proj_save: Optional[torch.Tensor] = None


_seed_pool: List[int] = []
_seed_state = None


def new_seed() -> int:
    """Dropout seed of one step, drawn from torch's CPU generator (deterministic under torch.manual_seed).
    Seeds are drawn 64 at a time; re-seeding the generator (a different initial_seed) discards the pooled ones."""
    global _seed_state
    st = torch.initial_seed()
    if not _seed_pool or _seed_state != st:
        _seed_state = st
        _seed_pool[:] = torch.randint(0, 2 ** 62, (64,), dtype=torch.int64).tolist()[::-1]
    return _seed_pool.pop()


# --------------------------------------------------------------------------------------------------
# phase runners (cut at BatchNorm statistics when they are shared across ranks)
# --------------------------------------------------------------------------------------------------
def _sync(world: int, opts: RunOptions) -> bool:
    return world > 1 and opts.sync_bn


def _run_head_forward(cs: CallState, plan: Plan, d, params_arr, tokens, mask, save, ws, out, attn):
    lib = L.lib()
    rm = L.ptr_array(cs.bn_running) if cs.bn_running else None
    tr = L.ptr_array(cs.bn_tracked) if cs.bn_tracked else None

    def call(p0, p1):
        L.check(lib.mvf_head_forward(C.byref(d), params_arr, rm, tr, L.ptr(tokens), L.ptr(mask), L.ptr(save),
                                     save.numel(), L.ptr(ws), ws.numel(), L.ptr(out), L.ptr(attn), p0, p1, _stream()),
                "mvf_head_forward")

    if _sync(plan.world, cs.opts) and cs.training:
        for ph in range(plan.n_fc + 1):
            call(ph, ph + 1)
            if ph < plan.n_fc:
                parallel.sync_stats_(plan.bn_stat(save, ph, False), cs.opts.process_group)
    else:
        call(0, L.PHASE_ALL)


def _run_head_backward(cs: CallState, plan: Plan, d, params_arr, tokens, mask, d_emb, save, ws, gpack, before_pool=None):
    """Backward phases 0 .. n_fc (cut at the BatchNorm statistics when they are shared across ranks), then the pooling
    backward as phase n_fc + 1; `before_pool` runs between the two (the chain's gradients are final at that point)."""
    lib = L.lib()

    def call(p0, p1):
        L.check(lib.mvf_head_backward(C.byref(d), params_arr, L.ptr(tokens), L.ptr(mask), L.ptr(d_emb), L.ptr(save),
                                      save.numel(), L.ptr(ws), ws.numel(), L.ptr(gpack), p0, p1, _stream()),
                "mvf_head_backward")

    if _sync(plan.world, cs.opts):
        for ph in range(plan.n_fc + 1):
            call(ph, ph + 1)
            if ph < plan.n_fc:
                parallel.sync_stats_(plan.bn_stat(save, plan.n_fc - 1 - ph, True), cs.opts.process_group)
    elif before_pool is not None:
        call(0, plan.n_fc + 1)
    else:
        call(0, L.PHASE_ALL)
        return
    if before_pool is not None:
        before_pool()
    call(plan.n_fc + 1, plan.n_fc + 2)


_comm_streams: Dict[torch.device, torch.cuda.Stream] = {}


def _comm_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _comm_streams.get(dev)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _comm_streams[dev] = s
    return s


def _run_proj_forward(cs: CallState, plan: Plan, d, params_arr, emb, save, ws, out):
    lib = L.lib()
    rm = L.ptr_array(cs.bn_running) if cs.bn_running else None
    tr = L.ptr_array(cs.bn_tracked) if cs.bn_tracked else None

    def call(p0, p1):
        L.check(lib.mvf_proj_forward(C.byref(d), params_arr, rm, tr, L.ptr(emb), int(cs.project), L.ptr(save),
                                     save.numel(), L.ptr(ws), ws.numel(), L.ptr(out), p0, p1, _stream()),
                "mvf_proj_forward")

    if _sync(plan.world, cs.opts) and cs.training and cs.project:
        call(0, 1)
        parallel.sync_stats_(plan.bn_stat(save, plan.n_fc, False), cs.opts.process_group)
        call(1, 2)
    else:
        call(0, L.PHASE_ALL)


def _run_proj_backward(cs: CallState, plan: Plan, d, params_arr, d_out, save, ws, gpack, d_emb):
    lib = L.lib()

    def call(p0, p1):
        L.check(lib.mvf_proj_backward(C.byref(d), params_arr, L.ptr(d_out), int(cs.project), L.ptr(save),
                                      save.numel(), L.ptr(ws), ws.numel(), L.ptr(gpack), L.ptr(d_emb), p0, p1, _stream()),
                "mvf_proj_backward")

    if _sync(plan.world, cs.opts) and cs.project:
        call(0, 1)
        parallel.sync_stats_(plan.bn_stat(save, plan.n_fc, True), cs.opts.process_group)
        call(1, 2)
    else:
        call(0, L.PHASE_ALL)


def _finish_grads(cs: CallState, plan: Plan, d, gpack: torch.Tensor, params: Sequence[Optional[torch.Tensor]],
                  reduced_from: Optional[int] = None):
    """One all-reduce of the flat gradient buffer (SUM), then scatter * 1/world into per-parameter tensors.
    The per-parameter gradients are views of ONE dense buffer (one allocation per step instead of one per parameter).
    reduced_from: elements [reduced_from:] were already all-reduced (overlapped with the pooling backward)."""
    if _world(cs.opts) > 1 and cs.opts.allreduce_grads:
        scale = parallel.finish_flat_grads_(gpack if reduced_from is None else gpack[:reduced_from], cs.opts.process_group)
    else:
        scale = 1.0
    present = [p for p in params if p is not None]
    if not present:
        return [None] * len(params)
    flat = torch.empty(sum(p.numel() for p in present), dtype=torch.float32, device=gpack.device)
    views = iter(torch._utils._unflatten_dense_tensors(flat, present))
    grads = [None if p is None else next(views) for p in params]
    arr = (C.c_void_p * len(params))()
    base, off = flat.data_ptr(), 0
    for i, p in enumerate(params):
        if p is not None:
            arr[i] = base + 4 * off
            off += p.numel()
    L.check(L.lib().mvf_unpack_grads(C.byref(d), L.ptr(gpack), arr, scale, _stream()), "mvf_unpack_grads")
    return grads


def _prep_tokens(tokens: torch.Tensor) -> torch.Tensor:
    _require_cuda(tokens, "tokens")
    if tokens.dtype == torch.float16:
        tokens = tokens.to(torch.bfloat16)
    if tokens.dim() != 4:
        raise ValueError(f"tokens must be [BV, T, P, C] token-major, got shape {tuple(tokens.shape)}")
    return tokens.contiguous()


def _prep_cls(cs: "CallState", frames: int) -> Optional[torch.Tensor]:
    """FIXED_WIDTH_BASELINE input: contiguous fp32 [frames, cls_dim] on the device (kept on the call state for backward)."""
    if cs.spec.pool_kind != "fwb":
        return None
    if cs.cls_emb is None:
        raise ValueError("FIXED_WIDTH_BASELINE: the head needs cls_emb (the CLS embedding of every frame)")
    c = cs.cls_emb.detach().reshape(frames, cs.spec.cls_dim).float().contiguous()
    _require_cuda(c, "cls_emb")
    cs.cls_emb = c
    return c


def _prep_mask(mask: Optional[torch.Tensor], BV: int, T: int, device) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    return mask.reshape(BV, T).to(device=device, dtype=torch.float32).contiguous()


# --------------------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, mask, cs: CallState, *params):
        tokens = _prep_tokens(tokens)
        BV, T, P, Cin = tokens.shape
        mask = _prep_mask(mask, BV, T, tokens.device)
        with torch.cuda.device(tokens.device):
            plan = Plan.get(cs.spec, BV, T, P, _mvf_dtype(tokens), cs.training, mask is not None, _bn_world(cs.opts),
                            cs.opts.gemm_backend, cs.opts.pool_mode)
            d = plan.desc_with_seed(cs.seed, cs.seed_dev, _prep_cls(cs, BV * T))
            save = torch.empty(plan.save_bytes, dtype=torch.uint8, device=tokens.device)
            ws = _scratch(tokens.device, plan.ws_bytes, "head")
            out = torch.empty(BV, T, cs.spec.emb, dtype=torch.float32, device=tokens.device)
            full = list(params) + [None] * (len(plan.param_names) - len(params))
            _run_head_forward(cs, plan, d, L.ptr_array(full), tokens, mask, save, ws, out, None)
        cs.plan, cs.head_save = plan, save
        ctx.cs, ctx.seed = cs, cs.seed
        ctx.n_params = len(params)
        ctx.save_for_backward(tokens, mask if mask is not None else torch.empty(0, device=tokens.device), *params)
        ctx.has_mask = mask is not None
        return out

    @staticmethod
    def backward(ctx, d_emb):
        cs: CallState = ctx.cs
        tokens, mask, *params = ctx.saved_tensors
        mask = mask if ctx.has_mask else None
        plan = cs.plan
        with torch.cuda.device(tokens.device):
            d = plan.desc_with_seed(ctx.seed, cs.seed_dev, cs.cls_emb)
            ws = _scratch(tokens.device, plan.ws_bytes, "head")
            gpack = parallel.flat_grad_buffer(plan.gpack_elems, tokens.device, cs.opts.process_group,
                                              enabled=_world(cs.opts) > 1 and cs.opts.allreduce_grads)
            full = list(params) + [None] * (len(plan.param_names) - len(params))
            _run_head_backward(cs, plan, d, L.ptr_array(full), tokens, mask, d_emb.contiguous().float(), cs.head_save,
                               ws, gpack)
            grads = _finish_grads(cs, plan, d, gpack, full)
        return (None, None, None) + tuple(grads[:ctx.n_params])


class ProjFn(torch.autograd.Function):
    """params: the six ssl_projection tensors in canonical order (ignored when cs.project is False)."""

    @staticmethod
    def forward(ctx, emb, cs: CallState, *params):
        _require_cuda(emb, "embeddings")
        BV, T, D = emb.shape
        embc = emb.contiguous().float()
        with torch.cuda.device(emb.device):
            plan = Plan.get(cs.spec, BV, T, 1, L.MVF_F32 if cs.opts.gemm_backend != L.GEMM_TCGEN05 else L.MVF_BF16,
                            cs.training, False, _bn_world(cs.opts), cs.opts.gemm_backend, cs.opts.pool_mode)
            d = plan.desc_with_seed(0)
            save = torch.empty(plan.proj_save_bytes, dtype=torch.uint8, device=emb.device)
            ws = _scratch(emb.device, plan.proj_ws_bytes, "proj")
            out = torch.empty(BV, T, D, dtype=torch.float32, device=emb.device)
            full = [None] * plan.n_head_params + list(params)
            if not cs.project:
                full = [None] * len(plan.param_names)
            _run_proj_forward(cs, plan, d, L.ptr_array(full), embc, save, ws, out)
        cs.plan, cs.proj_save = plan, save
        ctx.cs = cs
        ctx.n_params = len(params)
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        cs: CallState = ctx.cs
        params = list(ctx.saved_tensors)
        plan = cs.plan
        dev = d_out.device
        with torch.cuda.device(dev):
            d = plan.desc_with_seed(0)
            ws = _scratch(dev, plan.proj_ws_bytes, "proj")
            gpack = parallel.flat_grad_buffer(plan.gpack_elems, dev, cs.opts.process_group,
                                              enabled=_world(cs.opts) > 1 and cs.opts.allreduce_grads)
            d_emb = torch.empty(d_out.shape, dtype=torch.float32, device=dev)
            full = ([None] * plan.n_head_params + params) if cs.project else [None] * len(plan.param_names)
            _run_proj_backward(cs, plan, d, L.ptr_array(full), d_out.contiguous().float(), cs.proj_save, ws, gpack, d_emb)
            grads = _finish_grads(cs, plan, d, gpack, full) if cs.project else []
        pg = tuple(grads[plan.n_head_params:]) if cs.project else tuple([None] * ctx.n_params)
        return (d_emb, None) + pg


class ModelFn(torch.autograd.Function):
    """Head + projection (+normalise) as ONE autograd node: one flat gradient buffer, one all-reduce."""

    @staticmethod
    def forward(ctx, tokens, mask, cs: CallState, *params):
        tokens = _prep_tokens(tokens)
        BV, T, P, Cin = tokens.shape
        mask = _prep_mask(mask, BV, T, tokens.device)
        dev = tokens.device
        with torch.cuda.device(dev):
            plan = Plan.get(cs.spec, BV, T, P, _mvf_dtype(tokens), cs.training, mask is not None, _bn_world(cs.opts),
                            cs.opts.gemm_backend, cs.opts.pool_mode)
            d = plan.desc_with_seed(cs.seed, cs.seed_dev, _prep_cls(cs, BV * T))
            save = torch.empty(plan.save_bytes, dtype=torch.uint8, device=dev)
            psave = torch.empty(plan.proj_save_bytes, dtype=torch.uint8, device=dev)
            ws = _scratch(dev, plan.ws_bytes, "head")
            pws = _scratch(dev, plan.proj_ws_bytes, "proj")
            emb = torch.empty(BV, T, cs.spec.emb, dtype=torch.float32, device=dev)
            out = torch.empty(BV, T, cs.spec.emb, dtype=torch.float32, device=dev)
            arr = L.ptr_array(list(params))
            _run_head_forward(cs, plan, d, arr, tokens, mask, save, ws, emb, None)
            _run_proj_forward(cs, plan, d, arr, emb, psave, pws, out)
        cs.plan, cs.head_save, cs.proj_save = plan, save, psave
        ctx.cs, ctx.seed = cs, cs.seed
        # the parameters are only read through raw pointers in backward (before any optimizer step), so they are kept as
        # plain references: routing ~56 tensors through save_for_backward costs more host time than the SCL kernels
        ctx.tokens, ctx.mask, ctx.params, ctx.param_ptrs = tokens, mask, params, arr
        return out

    @staticmethod
    def backward(ctx, d_out):
        cs: CallState = ctx.cs
        tokens, mask, params = ctx.tokens, ctx.mask, ctx.params
        plan = cs.plan
        dev = tokens.device
        with torch.cuda.device(dev):
            d = plan.desc_with_seed(ctx.seed, cs.seed_dev, cs.cls_emb)
            ws = _scratch(dev, plan.ws_bytes, "head")
            pws = _scratch(dev, plan.proj_ws_bytes, "proj")
            gpack = parallel.flat_grad_buffer(plan.gpack_elems, dev, cs.opts.process_group,
                                              enabled=_world(cs.opts) > 1 and cs.opts.allreduce_grads)
            d_emb = torch.empty(d_out.shape, dtype=torch.float32, device=dev)
            arr = ctx.param_ptrs
            _run_proj_backward(cs, plan, d, arr, d_out.contiguous().float(), cs.proj_save, pws, gpack, d_emb)
            # several ranks, symmetric-memory buffer: the chain's gradients (12 of 19 MB at the Penn shape, final before the
            # pooling backward starts) are summed by a few CTAs on a second stream beside that HBM-bound kernel; only the
            # pooling gradients are left for the end of the step
            split = (plan.gpack_pool_elems + 3) // 4 * 4
            peer = parallel.PeerFlatGrads.find(gpack) if (_world(cs.opts) > 1 and cs.opts.allreduce_grads) else None
            # (only on the NVSwitch multimem path: plain peer loads / stores from a few CTAs are slower than the kernel they hide behind)
            if peer is not None and peer[0].mc_ptr and cs.opts.overlap_grad_allreduce and 0 < split < plan.gpack_elems:
                cur, comm = torch.cuda.current_stream(dev), _comm_stream(dev)

                lib = L.lib()
                reserve = max(1, int(cs.opts.pool_bwd_reserve_sms))

                def start_chain_allreduce():
                    comm.wait_stream(cur)
                    with torch.cuda.stream(comm):
                        parallel.finish_flat_grads_(gpack[split:], cs.opts.process_group, channel=1, ctas=reserve)
                    lib.mvf_pool_bwd_reserve_sms(reserve)

                try:
                    _run_head_backward(cs, plan, d, arr, tokens, mask, d_emb, cs.head_save, ws, gpack,
                                       before_pool=start_chain_allreduce)
                finally:
                    lib.mvf_pool_bwd_reserve_sms(0)
                cur.wait_stream(comm)
                grads = _finish_grads(cs, plan, d, gpack, list(params), reduced_from=split)
            else:
                _run_head_backward(cs, plan, d, arr, tokens, mask, d_emb, cs.head_save, ws, gpack)
                grads = _finish_grads(cs, plan, d, gpack, list(params))
        return (None, None, None) + tuple(grads)


class SCLFn(torch.autograd.Function):
    """loss = SCL(embs); the gradient w.r.t. embs is produced by the same launch sequence as the loss."""

    @staticmethod
    def forward(ctx, embs, seq_lens, steps, masks, temperature, label_variance, negative_type, quirk):
        _require_cuda(embs, "embeddings")
        Bv, V, T, D = embs.shape
        if V != 2:
            raise ValueError("SCL expects two views per video")
        dev = embs.device
        e = embs.contiguous().float()
        sl = seq_lens.reshape(Bv, 2).to(device=dev, dtype=torch.int64).contiguous()
        st = steps.reshape(Bv, 2, T).to(device=dev, dtype=torch.int64).contiguous()
        mk = masks.reshape(Bv, 2, T).to(device=dev, dtype=torch.float32).contiguous()
        lib = L.lib()
        with torch.cuda.device(dev):
            nbytes = lib.mvf_scl_ws_bytes(Bv, T, D)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            need_grad = ctx.needs_input_grad[0]
            dE = torch.empty_like(e) if need_grad else None
            L.check(lib.mvf_scl_fwd_bwd(L.ptr(e), L.ptr(sl), L.ptr(st), L.ptr(mk), Bv, T, D, float(temperature),
                                        float(label_variance), L.NEG[negative_type], 1 if quirk else 0, L.ptr(loss),
                                        L.ptr(dE), L.ptr(ws), nbytes, _stream()), "mvf_scl_fwd_bwd")
        ctx.dE = dE
        return loss

    @staticmethod
    def backward(ctx, g):
        dE = ctx.dE
        if dE is None:
            return (None,) * 8
        return (dE * g,) + (None,) * 7
