"""CPU oracle for the MV-Former head + SCL hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (module-free) restatement of the reference algorithm, written against
plain tensors so that it can run in fp32 or fp64 on the host.  It is the *checker* for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  Nothing under ``video_rep_learning_b200/`` imports it, and the
product path raises if the CUDA library is missing rather than falling back to this code.

Parity pinning: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against the reference's own PyTorch modules, imported in the build
container through ``oracle/ref_shim.py``; the comparison script is ``tests/golden/make_golden.py``
and the resulting vectors are committed under ``tests/golden/``.

Reference citations are relative to ``/root/reference/CARL_MVF``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# configuration of the head (the subset of cfg.MODEL.EMBEDDER_MODEL that shapes the math)
# --------------------------------------------------------------------------------------------------
@dataclass
class HeadCfg:
    """Hyper-parameters read by models/mvformer.py:20-115 and models/resnet_c2d.py:112-120."""
    c_in: int = 2304            # MODEL.BASE_MODEL.OUT_CHANNEL (x3 for SMART_FEATS "a,b,c")
    n_entities: int = 3         # SMART_TOKENS
    pool_channels: int = 384    # SMART_POOL_CHANNELS (default 384, mvformer.py:23-27)
    fc_channels: Tuple[int, ...] = (512, 512)   # FC_LAYERS channels x CAPACITY_SCALAR
    hidden: int = 256           # HIDDEN_SIZE
    d_ff: int = 1024            # D_FF
    n_heads: int = 8            # NUM_HEADS
    n_layers: int = 3           # NUM_LAYERS
    emb: int = 128              # EMBEDDING_SIZE
    proj: int = 128             # MODEL.PROJECTION_SIZE (hidden width of MLPHead)
    one_hot: str = "pool"       # SMART_ONE_HOT: none | pool | enc
    final: str = "one"          # SMART_FINAL: max | one | avg | lin
    train_frames: int = 20      # TRAIN.NUM_FRAMES (positional-encoder training length)
    drop_p: float = 0.0         # FC_DROPOUT_RATE (oracle only supports externally supplied masks)
    ln_eps: float = 1e-5
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    pool_kind: str = "lstp"     # "lstp": LearnableTokenPooling (entity cross-attention); "fwb": FIXED_WIDTH_BASELINE (FWBPooling)
    cls_dim: int = 0            # fwb: width of the CLS embedding = OUT_CHANNEL / len(SMART_FEATS) (mvformer.py:441-447)


def param_shapes(cfg: HeadCfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict names/shapes of `embed.*` and `ssl_projection.*`.

    Follows the registration order of mvformer.py:64-109, utils.py:196-209 and resnet_c2d.py:117-120.
    """
    E, SPC, H = cfg.n_entities, cfg.pool_channels, cfg.hidden
    s: Dict[str, Tuple[int, ...]] = {}
    if cfg.pool_kind == "fwb":                                   # FWBPooling.lin_conv (mvformer.py:449-451)
        s["embed.pooling.lin_conv.weight"] = (SPC * E, cfg.cls_dim)
        s["embed.pooling.lin_conv.bias"] = (SPC * E,)
    else:
        p = "embed.pooling.cross_att."
        s[p + "Q_s"] = (1, E, SPC)
        s[p + "Q_s_b"] = (SPC,)
        s[p + "linear_K2d.weight"] = (SPC, cfg.c_in)
        s[p + "linear_K2d.bias"] = (SPC,)
        s[p + "linear_V2d.weight"] = (SPC, cfg.c_in)
        s[p + "linear_V2d.bias"] = (SPC,)
    cin = SPC + (E if cfg.one_hot == "pool" else 0)
    for i, ch in enumerate(cfg.fc_channels):
        lin, bn = 4 * i + 1, 4 * i + 2
        s[f"embed.fc_layers.{lin}.weight"] = (ch, cin)
        s[f"embed.fc_layers.{lin}.bias"] = (ch,)
        s[f"embed.fc_layers.{bn}.weight"] = (ch,)
        s[f"embed.fc_layers.{bn}.bias"] = (ch,)
        cin = ch
    h_in = H - (E if cfg.one_hot == "enc" else 0)
    s["embed.video_emb.weight"] = (h_in, cin)
    s["embed.video_emb.bias"] = (h_in,)
    for l in range(cfg.n_layers):
        q = f"embed.video_encoder.enc_layers.{l}."
        s[q + "res_layer0.norm.weight"] = (H,)
        s[q + "res_layer0.norm.bias"] = (H,)
        s[q + "res_layer1.norm.weight"] = (H,)
        s[q + "res_layer1.norm.bias"] = (H,)
        for nm in ("linear_Q2d", "linear_K2d", "linear_V2d", "linear_d2Q"):
            s[q + f"self_att.{nm}.weight"] = (H, H)
            s[q + f"self_att.{nm}.bias"] = (H,)
        s[q + "feed_forward.fc1.weight"] = (cfg.d_ff, H)
        s[q + "feed_forward.fc1.bias"] = (cfg.d_ff,)
        s[q + "feed_forward.fc2.weight"] = (H, cfg.d_ff)
        s[q + "feed_forward.fc2.bias"] = (H,)
    s["embed.embedding_layer.weight"] = (cfg.emb, H)
    s["embed.embedding_layer.bias"] = (cfg.emb,)
    if cfg.final == "lin":
        s["embed.lin_final.weight"] = (H, E * H)
        s["embed.lin_final.bias"] = (H,)
    s["ssl_projection.net.0.weight"] = (cfg.proj, cfg.emb)
    s["ssl_projection.net.0.bias"] = (cfg.proj,)
    s["ssl_projection.net.1.weight"] = (cfg.proj,)
    s["ssl_projection.net.1.bias"] = (cfg.proj,)
    s["ssl_projection.net.3.weight"] = (cfg.emb, cfg.proj)
    s["ssl_projection.net.3.bias"] = (cfg.emb,)
    return s


def bn_buffer_names(cfg: HeadCfg) -> List[str]:
    """BatchNorm prefixes that carry running_mean / running_var / num_batches_tracked."""
    return [f"embed.fc_layers.{4 * i + 2}" for i in range(len(cfg.fc_channels))] + ["ssl_projection.net.1"]


def init_params(cfg: HeadCfg, seed: int = 1, dtype=torch.float32, scale: float = 1.0) -> Dict[str, Tensor]:
    """Deterministic synthetic parameters (NOT the reference's init; used by tests / bench / goldens).

    Matrices ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear; norm gains near 1; small biases.
    Every tensor is drawn from one CPU generator in `param_shapes` order so any host reproduces it.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shp in param_shapes(cfg).items():
        if name.endswith("norm.weight") or (name.endswith(".weight") and len(shp) == 1):
            t = 1.0 + 0.1 * (torch.rand(shp, generator=g, dtype=torch.float64) - 0.5)
        elif len(shp) == 1:
            t = 0.1 * (torch.rand(shp, generator=g, dtype=torch.float64) - 0.5)
        else:
            fan_in = shp[-1]
            t = (2.0 * torch.rand(shp, generator=g, dtype=torch.float64) - 1.0) * (scale / math.sqrt(fan_in))
        out[name] = t.to(dtype)
    return out


def init_bn_buffers(cfg: HeadCfg, dtype=torch.float32) -> Dict[str, Tensor]:
    buf: Dict[str, Tensor] = {}
    chans = list(cfg.fc_channels) + [cfg.proj]
    for pre, ch in zip(bn_buffer_names(cfg), chans):
        buf[pre + ".running_mean"] = torch.zeros(ch, dtype=dtype)
        buf[pre + ".running_var"] = torch.ones(ch, dtype=dtype)
        buf[pre + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return buf


# --------------------------------------------------------------------------------------------------
# a7: sin/cos table  (models/utils.py:113-126)
# --------------------------------------------------------------------------------------------------
def sincos_table(seq_len: int, d_model: int, train_len: Optional[int] = None) -> np.ndarray:
    """float64 [seq_len, d_model] table.

    The reference puts sin on EVEN channel indices and cos on ODD ones (its arrays are named the other
    way round, utils.py:114-115) and uses the channel index i itself in 10000**(i/d_model), not
    2*floor(i/2).  Positions are 0..S-1, or linspace(0, train_len-1, S) when S differs from the
    training length (utils.py:117-120, 138-143).
    """
    ch = np.arange(d_model, dtype=np.float64)
    if train_len is None:
        pos = np.arange(seq_len, dtype=np.float64)
    else:
        pos = np.linspace(0, train_len - 1, num=seq_len)
    ang = pos[:, None] / np.power(10000.0, ch[None, :] / d_model)
    tab = np.where((np.arange(d_model) % 2 == 0)[None, :], np.sin(ang), np.cos(ang))
    return tab


def pos_table_for(cfg: HeadCfg, S: int, d_model: int) -> np.ndarray:
    """PositionalEncoder.forward branch selection (models/utils.py:136-143)."""
    return sincos_table(S, d_model, None if S == cfg.train_frames else cfg.train_frames)


# --------------------------------------------------------------------------------------------------
# a3-a5: entity-query cross-attention pooling  (mvformer.py:243-266, 352-414; utils.py:11-44)
# --------------------------------------------------------------------------------------------------
def xattn_pool(P: Dict[str, Tensor], tokens: Tensor, cfg: HeadCfg) -> Tuple[Tensor, Tensor]:
    """tokens [BV,T,Ptok,C_in] (token-major) -> (ent [BV,T,E,SPC], attn [BV,T,E,Ptok]).

    K = X Wk^T + bk, V = X Wv^T + bv (mvformer.py:361-362); Q = Q_s + Q_s_b (mvformer.py:383);
    single head of width SPC, scores / sqrt(SPC), softmax over the patch axis, A @ V (utils.py:15-35).
    The reference loops over videos (mvformer.py:255-264) -- the result is independent per frame.
    """
    pre = "embed.pooling.cross_att."
    K = F.linear(tokens, P[pre + "linear_K2d.weight"], P[pre + "linear_K2d.bias"])
    V = F.linear(tokens, P[pre + "linear_V2d.weight"], P[pre + "linear_V2d.bias"])
    Q = P[pre + "Q_s"][0] + P[pre + "Q_s_b"]                       # [E, SPC]
    scores = torch.einsum("btpc,ec->btep", K, Q) / np.sqrt(cfg.pool_channels)
    attn = torch.softmax(scores, dim=-1)
    ent = torch.einsum("btep,btpc->btec", attn, V)
    return ent, attn


def fwb_pool(P: Dict[str, Tensor], cls_emb: Tensor, BV: int, T: int, cfg: HeadCfg) -> Tensor:
    """FIXED_WIDTH_BASELINE (FWBPooling.forward, mvformer.py:455-462): the patch tokens are ignored; one Linear maps the
    CLS embedding of every frame [BV*T, cls_dim] to SPC*E channels, reshaped [frames, SPC, E] -- "entity" e of channel c
    is output column c*E + e.  Returns ent [BV,T,E,SPC]."""
    y = F.linear(cls_emb, P["embed.pooling.lin_conv.weight"], P["embed.pooling.lin_conv.bias"])
    return y.reshape(BV, T, cfg.pool_channels, cfg.n_entities).permute(0, 1, 3, 2)


# --------------------------------------------------------------------------------------------------
# a6: per-entity MLP (mvformer.py:70-86, 144-153)
# --------------------------------------------------------------------------------------------------
def _batch_norm_train(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """Train-mode BatchNorm1d over rows: biased variance for normalisation. Returns (y, mean, var_b)."""
    mean = x.mean(dim=0)
    var = x.var(dim=0, unbiased=False)
    y = (x - mean) / torch.sqrt(var + eps) * w + b
    return y, mean, var


# Test hook: a callable (x, weight, bias, eps) -> (y, mean, biased_var) replacing the train-mode BatchNorm, used by
# the gloo tests to plug in a cross-rank (SyncBatchNorm-semantics) implementation.
BN_TRAIN_HOOK = None


def _bn(x, P, buf, pre, cfg: HeadCfg, training: bool, new_buf: Dict[str, Tensor]):
    if training:
        fn = BN_TRAIN_HOOK or _batch_norm_train
        y, mean, var = fn(x, P[pre + ".weight"], P[pre + ".bias"], cfg.bn_eps)
        if buf is not None:
            n = x.shape[0]
            m = cfg.bn_momentum
            unbiased = var * (n / max(n - 1, 1))
            new_buf[pre + ".running_mean"] = (1 - m) * buf[pre + ".running_mean"] + m * mean.detach()
            new_buf[pre + ".running_var"] = (1 - m) * buf[pre + ".running_var"] + m * unbiased.detach()
            new_buf[pre + ".num_batches_tracked"] = buf[pre + ".num_batches_tracked"] + 1
        return y
    rm, rv = buf[pre + ".running_mean"], buf[pre + ".running_var"]
    return (x - rm) / torch.sqrt(rv + cfg.bn_eps) * P[pre + ".weight"] + P[pre + ".bias"]


def entity_mlp(P, buf, ent: Tensor, cfg: HeadCfg, training: bool, new_buf, drop_masks=None) -> Tensor:
    """ent [BV,T,E,SPC] -> h3 [BV,T,E,H_in]; rows ordered (b,t,e) as in mvformer.py:151."""
    BV, T, E, SPC = ent.shape
    x = ent
    if cfg.one_hot == "pool":                                   # mvformer.py:144-149
        eye = torch.eye(E, dtype=ent.dtype, device=ent.device).expand(BV, T, E, E)
        x = torch.cat([x, eye], dim=-1)
    x = x.reshape(BV * T * E, x.shape[-1])
    for i in range(len(cfg.fc_channels)):
        lin, bn = 4 * i + 1, 4 * i + 2
        if drop_masks is not None:
            x = x * drop_masks[f"fc{i}"]
        x = F.linear(x, P[f"embed.fc_layers.{lin}.weight"], P[f"embed.fc_layers.{lin}.bias"])
        x = _bn(x, P, buf, f"embed.fc_layers.{bn}", cfg, training, new_buf)
        x = torch.relu(x)
    x = F.linear(x, P["embed.video_emb.weight"], P["embed.video_emb.bias"])   # mvformer.py:153
    return x.reshape(BV, T, E, x.shape[-1])


# --------------------------------------------------------------------------------------------------
# a8: temporal encoder (models/utils.py:47-108, 147-159, 176-242)
# --------------------------------------------------------------------------------------------------
ATTN_LEAN_ELEMS = 1 << 27   # score elements above which the oracle's attention runs view by view (memory only)


def _attention_core(q: Tensor, k: Tensor, v: Tensor, keymask: Optional[Tensor], dk: int) -> Tensor:
    """attention() of models/utils.py:11-44 with the [B,1,1,S] key mask, dropout p = 0."""
    sc = q @ k.transpose(-1, -2) / np.sqrt(dk)                                   # utils.py:17-18
    if keymask is not None:
        sc = sc.masked_fill(keymask[:, None, None, :] == 0, -float("inf"))      # utils.py:20-21
    return torch.softmax(sc, dim=-1) @ v


def encoder_layer(P, pre: str, z: Tensor, keymask: Optional[Tensor], cfg: HeadCfg, drop=None) -> Tensor:
    """Pre-LN block: z + MHA(LN z); z + FFN(LN z).  keymask [BV,S] (1 = attend, 0 = -inf)."""
    BV, S, H = z.shape
    nh = cfg.n_heads
    dk = H // nh
    r = F.layer_norm(z, (H,), P[pre + "res_layer0.norm.weight"], P[pre + "res_layer0.norm.bias"], cfg.ln_eps)
    a = pre + "self_att."
    q = F.linear(r, P[a + "linear_Q2d.weight"], P[a + "linear_Q2d.bias"]).view(BV, S, nh, dk).transpose(1, 2)
    k = F.linear(r, P[a + "linear_K2d.weight"], P[a + "linear_K2d.bias"]).view(BV, S, nh, dk).transpose(1, 2)
    v = F.linear(r, P[a + "linear_V2d.weight"], P[a + "linear_V2d.bias"]).view(BV, S, nh, dk).transpose(1, 2)
    if BV * nh * S * S > ATTN_LEAN_ELEMS and BV > 1:
        # long sequences (S = E*T up to thousands): the [BV, heads, S, S] score tensors of all views and layers would be
        # kept for backward; evaluate view by view under activation checkpointing instead (same arithmetic)
        from torch.utils.checkpoint import checkpoint
        ctx = torch.cat([checkpoint(_attention_core, q[b:b + 1], k[b:b + 1], v[b:b + 1],
                                    None if keymask is None else keymask[b:b + 1], dk, use_reentrant=False)
                         for b in range(BV)], dim=0)
    else:
        ctx = _attention_core(q, k, v, keymask, dk)
    ctx = ctx.transpose(1, 2).reshape(BV, S, H)
    o = F.linear(ctx, P[a + "linear_d2Q.weight"], P[a + "linear_d2Q.bias"])
    if drop is not None:
        o = o * drop[0]
    z = z + o
    r = F.layer_norm(z, (H,), P[pre + "res_layer1.norm.weight"], P[pre + "res_layer1.norm.bias"], cfg.ln_eps)
    f = torch.relu(F.linear(r, P[pre + "feed_forward.fc1.weight"], P[pre + "feed_forward.fc1.bias"]))
    g = F.linear(f, P[pre + "feed_forward.fc2.weight"], P[pre + "feed_forward.fc2.bias"])
    if drop is not None:
        g = g * drop[1]
    return z + g


# --------------------------------------------------------------------------------------------------
# a2 + a9: whole head  (mvformer.py:128-200)
# --------------------------------------------------------------------------------------------------
def head_forward(P: Dict[str, Tensor], buf: Optional[Dict[str, Tensor]], tokens: Tensor,
                 video_masks: Optional[Tensor], cfg: HeadCfg, training: bool = True,
                 drop_masks: Optional[Dict[str, Tensor]] = None,
                 return_aux: bool = False, cls_emb: Optional[Tensor] = None):
    """tokens [BV,T,Ptok,C_in] token-major; video_masks [BV,1,T] or [BV,T] or None -> emb [BV,T,D].

    Equivalent to MultiEntityTransformerEmbModel.forward on x = tokens.permute(0,1,3,2).reshape(BV,T,C,h,w).
    `drop_masks` (optional) holds pre-scaled keep masks: 'fc0','fc1' [R,c_in], 'pos' [BV*E,T,H],
    'enc{l}_0','enc{l}_1' [BV,S,H]; absent -> dropout is the identity (p = 0 / eval).
    Returns (emb, new_buffers[, aux]).
    """
    BV, T, Ptok, C = tokens.shape
    E, H = cfg.n_entities, cfg.hidden
    new_buf: Dict[str, Tensor] = {}
    if cfg.pool_kind == "fwb":
        ent, attn = fwb_pool(P, cls_emb, BV, T, cfg), None
    else:
        ent, attn = xattn_pool(P, tokens, cfg)
    h3 = entity_mlp(P, buf, ent, cfg, training, new_buf, drop_masks)             # [BV,T,E,Hin]
    z = h3.permute(0, 2, 1, 3)                                                   # [BV,E,T,Hin]  mvformer.py:155-157
    pe = torch.from_numpy(pos_table_for(cfg, T, z.shape[-1])).to(device=z.device, dtype=z.dtype)   # utils.py:136-143
    z = z + pe[None, None]
    if drop_masks is not None and "pos" in drop_masks:
        z = z * drop_masks["pos"].reshape(BV, E, T, -1)
    if cfg.one_hot == "enc":                                                     # mvformer.py:162-168
        eye = torch.eye(E, dtype=z.dtype, device=z.device)[None, :, None, :].expand(BV, E, T, E)
        z = torch.cat([z, eye], dim=-1)
    z = z.reshape(BV, E * T, z.shape[-1])                                        # s = e*T + t   mvformer.py:170
    keymask = None
    if video_masks is not None:                                                  # mvformer.py:174-177
        vm = video_masks.reshape(BV, T)
        keymask = vm[:, None, :].expand(BV, E, T).reshape(BV, E * T)
    for l in range(cfg.n_layers):
        dm = None
        if drop_masks is not None and f"enc{l}_0" in drop_masks:
            dm = (drop_masks[f"enc{l}_0"], drop_masks[f"enc{l}_1"])
        z = encoder_layer(P, f"embed.video_encoder.enc_layers.{l}.", z, keymask, cfg, dm)
    z4 = z.view(BV, E, T, H)
    if cfg.final == "max":                                                       # mvformer.py:182-195
        y = z4.max(dim=1)[0]
    elif cfg.final == "one":
        y = z4[:, 0]
    elif cfg.final == "avg":
        y = z4.mean(dim=1)
    elif cfg.final == "lin":
        y = F.linear(z4.permute(0, 2, 1, 3).reshape(BV, T, E * H), P["embed.lin_final.weight"], P["embed.lin_final.bias"])
    else:
        raise ValueError(cfg.final)
    emb = F.linear(y.reshape(BV * T, H), P["embed.embedding_layer.weight"], P["embed.embedding_layer.bias"])
    emb = emb.view(BV, T, cfg.emb)
    if return_aux:
        return emb, new_buf, {"ent": ent, "attn": attn, "h3": h3, "z": z}
    return emb, new_buf


# --------------------------------------------------------------------------------------------------
# a10: projection MLP + L2 normalise (resnet_c2d.py:112-126; transformer.py:226-230)
# --------------------------------------------------------------------------------------------------
def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


def proj_forward(P, buf, emb: Tensor, cfg: HeadCfg, training: bool = True):
    """emb [BV,T,D] -> unit-norm projected embeddings [BV,T,D]; hidden width = PROJECTION_SIZE."""
    BV, T, D = emb.shape
    new_buf: Dict[str, Tensor] = {}
    x = emb.reshape(BV * T, D)
    x = F.linear(x, P["ssl_projection.net.0.weight"], P["ssl_projection.net.0.bias"])
    x = _bn(x, P, buf, "ssl_projection.net.1", cfg, training, new_buf)
    x = torch.relu(x)
    x = F.linear(x, P["ssl_projection.net.3.weight"], P["ssl_projection.net.3.bias"])
    return l2_normalize(x).view(BV, T, D), new_buf


def model_forward(P, buf, tokens, video_masks, cfg: HeadCfg, project: bool, l2_norm: bool = True,
                  training: bool = True, drop_masks=None):
    """TransformerModel.forward after the backbone hand-off (transformer.py:219-230)."""
    emb, nb = head_forward(P, buf, tokens, video_masks, cfg, training, drop_masks)
    if project:
        out, nb2 = proj_forward(P, buf, emb, cfg, training)
        nb.update(nb2)
        return out, nb
    if l2_norm:
        return l2_normalize(emb), nb
    return emb, nb


# --------------------------------------------------------------------------------------------------
# a12: Sequence Contrastive Loss  (algos/scl.py:52-105)
# --------------------------------------------------------------------------------------------------
def scl_loss_dense(embs: Tensor, seq_lens: Tensor, steps: Tensor, masks: Tensor,
                   temperature: float = 0.1, label_variance: float = 10.0,
                   negative_type: str = "single_noself") -> Tensor:
    """Full N x N statement of the loss, N = Bv*2*T, vectorised over the batch.

    embs [Bv,2,T,D] (unit rows), seq_lens [Bv,2] int, steps [Bv,2,T] int, masks [Bv*2,1,T] float.
    Index bookkeeping replaces the reference's per-video slice loops (scl.py:68-79, 89-96):
      vid(i) = i // (2T), view(i) = (i // T) % 2.
    """
    Bv, V, T, D = embs.shape
    N = Bv * V * T
    e = embs.reshape(N, D)
    st = steps.reshape(N)
    L = seq_lens.reshape(Bv, V, 1).expand(Bv, V, T).reshape(N).to(torch.float32)   # scl.py:58 `.float()`
    m = masks.reshape(N).to(e.dtype)
    mm = m[:, None] * m[None, :]                                                 # scl.py:59
    logits = (e @ e.t()) / temperature                                           # scl.py:61
    # int64 / float32 -> float32: the timestamp distance and the Gaussian are evaluated in float32 whatever
    # the dtype of the embeddings, then cast with .type_as(logits) (scl.py:62, 85)
    dist = torch.abs(st[:, None] / L[:, None] * L[None, :] - st[None, :])        # scl.py:62
    dist = dist.masked_fill(mm == 0, 1e6)                                        # scl.py:63
    idx = torch.arange(N, device=e.device)
    vid, view = idx // (V * T), (idx // T) % V
    same_vid = vid[:, None] == vid[None, :]
    same_view = same_vid & (view[:, None] == view[None, :])
    w = torch.ones_like(logits)
    if "single" in negative_type:                                                # scl.py:74-76
        w = torch.where(same_vid, w, torch.zeros_like(w))
    if "noself" in negative_type:                                                # scl.py:77-79
        w = torch.where(same_view, torch.zeros_like(w), w)
    w = w.masked_fill(mm == 0, 1e-6)                                             # scl.py:80 (runs last)
    pos = torch.exp(-torch.square(dist) / (2 * label_variance)).to(e.dtype)      # scl.py:85
    cross = same_vid & ~same_view                                                # own video, other view
    pos = torch.where(cross, pos, torch.zeros_like(pos))
    den = pos.sum(dim=1, keepdim=True)
    label = torch.where(den > 0, pos / den.clamp_min(1e-300 if e.dtype == torch.float64 else 1e-38),
                        torch.zeros_like(pos))                                   # safe_div scl.py:13-16
    ex = torch.exp(logits)
    Z = (w * ex).sum(dim=1, keepdim=True)                                        # scl.py:98-99
    logq = torch.log(ex / Z + 1e-6)
    kl = torch.where(label > 0, label * (torch.log(label.clamp_min(1e-300 if e.dtype == torch.float64 else 1e-45)) - logq),
                     torch.zeros_like(label))                                    # F.kl_div pointwise
    return (kl * mm).sum() / m.sum()                                             # scl.py:102-103


def scl_pair_closed_form(E0: np.ndarray, E1: np.ndarray, s0: np.ndarray, s1: np.ndarray,
                         m0: np.ndarray, m1: np.ndarray, L0: float, L1: float, M: float,
                         temperature: float = 0.1, label_variance: float = 10.0,
                         extra_cols: Optional[np.ndarray] = None):
    """Per-video-pair statement of `single_noself` with its closed-form gradient (float64 numpy).

    One T x T matrix S = E0 E1^T / tau serves both view directions.
    `extra_cols` [U,D]: embeddings of every masked frame in the local batch (any video / view): each
    contributes weight 1e-6 to every valid row's partition sum (the scl.py:80 quirk).
    Returns loss contribution (already divided by M), dE0, dE1, and d(extra_cols).
    The timestamp arithmetic is done in float32 exactly as the reference does (scl.py:57-62).
    """
    T, D = E0.shape
    tau = temperature
    S = (E0 @ E1.T) / tau
    f32 = np.float32
    d01 = np.abs((s0.astype(f32)[:, None] / f32(L0)) * f32(L1) - s1.astype(f32)[None, :]).astype(np.float64)
    d10 = np.abs((s1.astype(f32)[:, None] / f32(L1)) * f32(L0) - s0.astype(f32)[None, :]).astype(np.float64)
    mm = m0[:, None] * m1[None, :]
    out = []
    U = None if extra_cols is None or len(extra_cols) == 0 else extra_cols
    for direction in (0, 1):
        Sd = S if direction == 0 else S.T
        dd = d01 if direction == 0 else d10
        mmd = mm if direction == 0 else mm.T
        Er = E0 if direction == 0 else E1
        dd = np.where(mmd == 0, f32(1e6), dd.astype(f32))
        # float32 Gaussian evaluated with torch's expf so that it is the reference's own rounding (scl.py:85)
        pw = torch.exp(torch.from_numpy(-(dd * dd) / f32(2 * label_variance))).numpy().astype(np.float64)
        den = pw.sum(1, keepdims=True)
        with np.errstate(invalid="ignore", divide="ignore"):
            y = np.where(den > 0, pw / den, 0.0)
        ex = np.exp(Sd)
        Z = (ex * mmd).sum(1)
        if U is not None:
            exU = np.exp(Er @ U.T / tau)
            Z = Z + 1e-6 * exU.sum(1)
        rowvalid = (m0 if direction == 0 else m1) > 0
        Zs = np.where(Z > 0, Z, 1.0)
        p = ex / Zs[:, None]
        q = p + 1e-6
        with np.errstate(invalid="ignore", divide="ignore"):
            kl = np.where(y > 0, y * (np.log(np.where(y > 0, y, 1.0)) - np.log(q)), 0.0)
        loss = (kl * mmd).sum() / M
        r = p / q
        g = (y * mmd * r).sum(1)
        G = (mmd * p * g[:, None] - y * mmd * r) / M                 # d loss / d Sd on pair entries
        G = np.where(rowvalid[:, None], G, 0.0)
        if U is not None:
            GU = np.where(rowvalid[:, None], 1e-6 * (exU / Zs[:, None]) * g[:, None] / M, 0.0)
        else:
            GU = None
        out.append((loss, G, GU))
    (l0, G0, GU0), (l1, G1, GU1) = out
    Gs = G0 + G1.T
    dE0 = Gs @ E1 / tau
    dE1 = Gs.T @ E0 / tau
    dU = None
    if U is not None:
        dE0 = dE0 + GU0 @ U / tau
        dE1 = dE1 + GU1 @ U / tau
        dU = (GU0.T @ E0 + GU1.T @ E1) / tau
    return l0 + l1, dE0, dE1, dU


def scl_loss_pairs(embs: np.ndarray, seq_lens: np.ndarray, steps: np.ndarray, masks: np.ndarray,
                   temperature: float = 0.1, label_variance: float = 10.0, quirk: bool = True):
    """Whole-batch `single_noself` loss + dE through the per-pair closed form (float64).

    embs [Bv,2,T,D]; masks [Bv,2,T].  With quirk=True every masked frame of the local batch joins
    every valid row's partition sum with weight 1e-6 (exactly scl.py:80); masked frames then also
    receive a (tiny) gradient.
    """
    Bv, V, T, D = embs.shape
    embs = embs.astype(np.float64)
    masks = masks.reshape(Bv, V, T).astype(np.float64)
    M = masks.sum()
    flatE = embs.reshape(-1, D)
    flatM = masks.reshape(-1)
    uidx = np.nonzero(flatM == 0)[0]
    U = flatE[uidx] if (quirk and len(uidx)) else None
    dE = np.zeros_like(embs)
    dflat = dE.reshape(-1, D)
    total = 0.0
    for v in range(Bv):
        l, d0, d1, dU = scl_pair_closed_form(embs[v, 0], embs[v, 1], steps[v, 0], steps[v, 1],
                                             masks[v, 0], masks[v, 1], float(seq_lens[v, 0]), float(seq_lens[v, 1]),
                                             M, temperature, label_variance, U)
        total += l
        dE[v, 0] += d0
        dE[v, 1] += d1
        if dU is not None:
            dflat[uidx] += dU
    return total, dE


# --------------------------------------------------------------------------------------------------
# a13: two-view temporal sampling, integer-exact (datasets/penn_action.py:152-206 and variants)
# --------------------------------------------------------------------------------------------------
def sample_frames_oracle(seq_len: int, num_frames: int, pre_steps=None, *, variant: str = "penn_action",
                         sampling_region: float = 1.5, consistent_offset: float = 0.2,
                         strategy: str = "time_augment"):
    """Restates the RNG call ORDER of the reference sampler: np.random.uniform -> np.random.randint ->
    torch.randperm.  `variant` selects the block-size rule: penn_action / kinetics400 / pouring use
    ceil(r*seq_len) (penn_action.py:170-172, kinetics400.py:135-182, pouring.py:150-154); finegym uses
    ceil(r*num_valid) (finegym.py:187); pouring with SAMPLE_FIX ('pouring_fix') uses ceil(r*num_frames).
    Returns (steps, chosen_steps, video_mask) for NUM_CONTEXTS == 1.
    """
    pre_offset = min(pre_steps) if pre_steps is not None else None
    if strategy == "offset_uniform":
        if seq_len >= num_frames:
            steps = torch.sort(torch.randperm(seq_len)[:num_frames])[0]
        else:
            steps = torch.arange(0, num_frames)
    elif strategy == "time_augment":
        num_valid = min(seq_len, num_frames)
        ratio = np.random.uniform(low=1.0, high=sampling_region) if sampling_region > 1 else 1.0
        base = {"finegym": num_valid, "pouring_fix": num_frames}.get(variant, seq_len)
        block = math.ceil(ratio * base)
        if pre_steps is not None and consistent_offset != 0:
            shift = int((1 - consistent_offset) * num_valid)
            lo = max(0, min(seq_len - block, pre_offset - shift))
            hi = max(1, min(seq_len - block + 1, pre_offset + shift + 1))
            offset = np.random.randint(low=lo, high=hi)
        else:
            offset = np.random.randint(low=0, high=max(seq_len - block, 1))
        steps = torch.sort(offset + torch.randperm(block)[:num_valid])[0]
        if num_valid < num_frames:
            steps = F.pad(steps, (0, num_frames - num_valid), "constant", seq_len)
    else:
        raise ValueError(strategy)
    mask = torch.ones(num_frames)
    mask[steps < 0] = 0
    mask[steps >= seq_len] = 0
    chosen = torch.clamp(steps.clone(), 0, seq_len - 1)
    return chosen.clone(), chosen, mask


# --------------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench / golden generation
# --------------------------------------------------------------------------------------------------
def synth_batch(Bv: int, T: int, Ptok: int, c_in: int, seed: int = 1, dtype=torch.float32,
                with_padding: bool = True):
    """Seeded synthetic step input: tokens [Bv*2,T,Ptok,C_in] ~ N(0,1); per-video seq_len in
    [ceil(T/2), 3T]; steps / masks from the sampler restatement (so padding and clamping occur)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    tokens = torch.randn(Bv * 2, T, Ptok, c_in, generator=g, dtype=torch.float32).to(dtype)
    rs_np, rs_t = np.random.get_state(), torch.random.get_rng_state()
    np.random.seed(seed)
    torch.manual_seed(seed)
    seq_lens = torch.zeros(Bv, 2, dtype=torch.int64)
    steps = torch.zeros(Bv, 2, T, dtype=torch.int64)
    masks = torch.ones(Bv, 2, T, dtype=torch.float32)
    for b in range(Bv):
        L = int(np.random.randint(math.ceil(T / 2), 3 * T + 1)) if with_padding else 4 * T
        seq_lens[b] = L
        _, c0, m0 = sample_frames_oracle(L, T, None)
        _, c1, m1 = sample_frames_oracle(L, T, c0)
        steps[b, 0], steps[b, 1] = c0, c1
        masks[b, 0], masks[b, 1] = m0, m1
        if masks[b].sum() == 0:                       # never feed an all-masked video
            masks[b, :, 0] = 1
    np.random.set_state(rs_np)
    torch.random.set_rng_state(rs_t)
    return tokens, seq_lens, steps, masks.reshape(Bv * 2, 1, T)
