"""MV-Former head (multi-entity fusion transformer) with the reference's module API, executed as fused
sm_100a kernels.

Mirrors CARL_MVF/models/mvformer.py: same class names, constructor arguments, attribute / parameter names
and registration order (so `state_dict()` is interchangeable and a seeded construction draws the same
initial values), same `forward(x, video_masks=None, cls_emb=None)` signature, `set_warmup_status`,
`embedding_size`, and the `pooling.cross_att.attn_holder` / `attn_matrix` side channel that
visualize_lstp.py hooks (mvformer.py:346-348, 408-411).

What differs is HOW forward runs: one call into libmvf_b200.so (engine.HeadFn) instead of a Python loop
over videos + ~100 ATen launches, and the input may be given token-major ([BV,T,P,C], the ViT's native
layout, no permute copy) as well as in the reference's NCHW form ([BV,T,C,h,w]).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from .. import engine
from .utils import Encoder, PositionalEncoder, _FusedOnly

_NEXT_ROUND = ("not implemented in this build (SURVEY.md section 8f, optional head branches); "
               "the shipped configs_mvf/{penn,fg99,fg288,pouring,k400}_mvf.yml do not use it")


def _get(cfg_node, key, default):
    return cfg_node[key] if key in cfg_node else default


def head_spec_from_cfg(cfg) -> engine.HeadSpec:
    """cfg keys -> HeadSpec (mvformer.py:20-115, resnet_c2d.py:115)."""
    em = cfg.MODEL.EMBEDDER_MODEL
    cap = em.CAPACITY_SCALAR
    fc = tuple(int(ch) * cap for ch, _act in em.FC_LAYERS) if "FC_LAYERS" in em and em.FC_LAYERS is not None else ()
    return engine.HeadSpec(
        c_in=int(cfg.MODEL.BASE_MODEL.OUT_CHANNEL),
        n_entities=int(_get(em, "SMART_TOKENS", 5)),
        pool_channels=int(_get(em, "SMART_POOL_CHANNELS", 384)),
        fc_channels=fc,
        hidden=int(em.HIDDEN_SIZE), d_ff=int(em.D_FF), n_heads=int(em.NUM_HEADS), n_layers=int(em.NUM_LAYERS),
        emb=int(em.EMBEDDING_SIZE),
        proj=int(_get(cfg.MODEL, "PROJECTION_SIZE", em.EMBEDDING_SIZE)),
        one_hot=str(_get(em, "SMART_ONE_HOT", "none")),
        final=str(_get(em, "SMART_FINAL", "max")),
        train_frames=int(cfg.TRAIN.NUM_FRAMES),
        drop_p=float(em.FC_DROPOUT_RATE),
        pool_kind="fwb" if _get(em, "FIXED_WIDTH_BASELINE", False) else "lstp",
        cls_dim=_cls_width(cfg) if _get(em, "FIXED_WIDTH_BASELINE", False) else 0,
    )


def _cls_width(cfg) -> int:
    """d_dyn_in of mvformer.py:226-232 / 441-447: OUT_CHANNEL divided by the number of SMART_FEATS layers."""
    em = cfg.MODEL.EMBEDDER_MODEL
    d = int(cfg.MODEL.BASE_MODEL.OUT_CHANNEL)
    if "SMART_FEATS" in em:
        sfl = str(em.SMART_FEATS)
        if "," in sfl:
            d = int(d / len(sfl.split(",")))
    return d


class LSTPCrossAtt(_FusedOnly):
    """Learnable static entity queries + K/V projections (mvformer.py:275-348)."""

    def __init__(self, cfg, num_static, num_dynamic, d_model_K, d_model_V, d_model, d_dyn_in=None, dout_p=0.0):
        super().__init__()
        self.cfg = cfg
        self.d_model_K, self.d_model_V, self.d_model = d_model_K, d_model_V, d_model
        em = cfg.MODEL.EMBEDDER_MODEL
        for key in ("VAL_PASS", "SMART_DISJOINT", "SMART_LN_KEYS"):
            if _get(em, key, False):
                raise NotImplementedError(f"MODEL.EMBEDDER_MODEL.{key} is {_NEXT_ROUND}")
        if num_dynamic > 0:
            raise NotImplementedError(f"SMART_DYNAMIC_TOKENS > 0 is {_NEXT_ROUND}")
        if num_static == 0:
            # same condition the reference exits on (mvformer.py:315-317)
            raise ValueError("cannot have both num_static == 0 and num_dynamic == 0")
        self.pass_through = False
        self.disjoint_att = False
        self.ln_keys = False
        self.dyn_ctrl = "separate"
        self.linear_K2d = nn.Linear(d_model_K, d_model)
        self.linear_V2d = nn.Linear(d_model_V, d_model)
        self.num_s, self.stat = num_static, True
        self.Q_s = nn.Parameter(torch.empty([1, num_static, d_model], dtype=torch.float32))
        nn.init.kaiming_uniform_(self.Q_s, a=math.sqrt(5))
        self.Q_s_b = nn.Parameter(torch.empty(d_model, dtype=torch.float32))
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.Q_s)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.Q_s_b, -bound, bound)
        self.num_d, self.dyn = 0, False
        self.dout_p = dout_p
        self.dropout = nn.Dropout(dout_p)
        self.visual = True
        self.attn_holder = nn.Identity()
        self.attn_matrix = None


class LearnableTokenPooling(_FusedOnly):
    """mvformer.py:207-238 -- owns `cross_att`."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        em = cfg.MODEL.EMBEDDER_MODEL
        self.nst = int(_get(em, "SMART_TOKENS", 5))
        self.nsdt = int(_get(em, "SMART_DYNAMIC_TOKENS", 0))
        self.spc = int(_get(em, "SMART_POOL_CHANNELS", 384))
        self.in_c = int(cfg.MODEL.BASE_MODEL.OUT_CHANNEL)
        d_dyn_in = self.in_c
        if "SMART_FEATS" in em:
            sfl = str(em.SMART_FEATS)
            if "," in sfl:
                d_dyn_in = int(d_dyn_in / len(sfl.split(",")))
        self.cross_att = LSTPCrossAtt(cfg=cfg, num_static=self.nst, num_dynamic=self.nsdt, d_model_K=self.in_c,
                                      d_model_V=self.in_c, d_model=self.spc, d_dyn_in=d_dyn_in)


class FWBPooling(_FusedOnly):
    """FIXED_WIDTH_BASELINE (mvformer.py:421-462): `lin_conv` maps the CLS embedding of a frame to SPC * tokens channels;
    parameter container only -- the product runs inside the fused head (csrc/head.cu, pool_kind FWB)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        em = cfg.MODEL.EMBEDDER_MODEL
        self.nst = int(_get(em, "SMART_TOKENS", 5))
        self.nsdt = int(_get(em, "SMART_DYNAMIC_TOKENS", 0))
        self.spc = int(_get(em, "SMART_POOL_CHANNELS", 384))
        self.in_c = int(cfg.MODEL.BASE_MODEL.OUT_CHANNEL)
        self.lin_conv = nn.Linear(_cls_width(cfg), self.spc * (self.nst + self.nsdt))


class MultiEntityTransformerEmbModel(nn.Module):
    """Drop-in for CARL_MVF/models/mvformer.py:15 (`model.embed`)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        em = cfg.MODEL.EMBEDDER_MODEL
        drop_rate = em.FC_DROPOUT_RATE
        in_channels = int(_get(em, "SMART_POOL_CHANNELS", 384))
        self.nst = int(_get(em, "SMART_TOKENS", 5))
        self.nsdt = int(_get(em, "SMART_DYNAMIC_TOKENS", 0))
        self.one_hot_pos = str(_get(em, "SMART_ONE_HOT", "none"))
        assert self.one_hot_pos in ["none", "pool", "enc"]
        if self.one_hot_pos == "enc":
            raise NotImplementedError(f"SMART_ONE_HOT: enc is {_NEXT_ROUND}")
        if self.one_hot_pos == "pool":
            in_channels += self.nst + self.nsdt
        self.fwb = bool(_get(em, "FIXED_WIDTH_BASELINE", False))
        if self.fwb and self.nsdt > 0:
            raise NotImplementedError(f"FIXED_WIDTH_BASELINE with SMART_DYNAMIC_TOKENS > 0 is {_NEXT_ROUND}")
        cap_scalar = em.CAPACITY_SCALAR
        fc_params = em.FC_LAYERS if "FC_LAYERS" in em else None
        self.embedding_size = em.EMBEDDING_SIZE
        hidden_channels = em.HIDDEN_SIZE

        self.pooling = FWBPooling(cfg) if self.fwb else LearnableTokenPooling(cfg)
        if fc_params is None:
            self.fc_layers = nn.Identity()
        else:
            layers: List[nn.Module] = []
            for channels, _activate in fc_params:
                channels = channels * cap_scalar
                layers += [nn.Dropout(drop_rate), nn.Linear(in_channels, channels), nn.BatchNorm1d(channels), nn.ReLU(True)]
                in_channels = channels
            self.fc_layers = nn.Sequential(*layers)
        self.video_emb = nn.Linear(in_channels, hidden_channels)
        self.video_pos_enc = PositionalEncoder(cfg, hidden_channels, drop_rate, seq_len=cfg.TRAIN.NUM_FRAMES)
        if em.NUM_LAYERS > 0:
            self.video_encoder = Encoder(hidden_channels, drop_rate, em.NUM_HEADS, em.D_FF, em.NUM_LAYERS)
        self.embedding_layer = nn.Linear(hidden_channels, self.embedding_size)
        self.smart_final = str(_get(em, "SMART_FINAL", "max"))
        assert self.smart_final in ["max", "one", "avg", "lin"]
        if self.smart_final == "lin":
            self.lin_final = nn.Linear((self.nst + self.nsdt) * hidden_channels, hidden_channels)
        self.in_backbone_warmup = "BACKBONE_WARMUP" in cfg.TRAIN

        self.spec = head_spec_from_cfg(cfg)
        self.run_options = engine.RunOptions()
        self._param_order: Optional[List[str]] = None
        self._param_slots = None
        self.last_call: Optional[engine.CallState] = None
        self.seed_dev: Optional[torch.Tensor] = None   # set by graph.GraphedTrainStep: device-side dropout counter

    # ---- reference API -------------------------------------------------------------------------------
    def set_warmup_status(self, new_status):
        # tokens are consumed without a gradient path in this build (frozen backbone), so warm-up is a flag only
        self.in_backbone_warmup = new_status

    # ---- plumbing --------------------------------------------------------------------------------------
    def head_param_names(self) -> List[str]:
        """Canonical (C ABI) order of this module's parameters, without the 'embed.' prefix."""
        if self._param_order is None:
            plan = engine.Plan.get(self.spec, 1, 1, 1, 0, False, False, 1, 0)
            self._param_order = [n[len("embed."):] for n in plan.param_names if n.startswith("embed.")]
        return self._param_order

    def head_params(self) -> List[torch.Tensor]:
        # (owning module, attribute) pairs are resolved once; the per-step cost is one dict lookup per parameter,
        # and a re-assigned parameter (module.weight = nn.Parameter(...)) is still picked up
        if self._param_slots is None:
            slots = []
            for n in self.head_param_names():
                mod_path, _, attr = n.rpartition(".")
                slots.append((self.get_submodule(mod_path)._parameters if mod_path else self._parameters, attr))
            self._param_slots = slots
        return [d[a] for d, a in self._param_slots]

    def bn_buffers(self):
        bns = [m for m in self.fc_layers if isinstance(m, nn.BatchNorm1d)] if isinstance(self.fc_layers, nn.Sequential) else []
        running, tracked = [], []
        for bn in bns:
            running += [bn.running_mean, bn.running_var]
            tracked.append(bn.num_batches_tracked)
        return running, tracked

    @staticmethod
    def to_token_major(x: torch.Tensor) -> torch.Tensor:
        """Accept the reference layout [BV,T,C,h,w] (transformer.py:203-214) or token-major [BV,T,P,C]."""
        if x.dim() == 5:
            BV, T, Cc, h, w = x.shape
            return x.reshape(BV, T, Cc, h * w).transpose(2, 3).contiguous()
        if x.dim() == 4:
            return x
        raise ValueError(f"expected [BV,T,C,h,w] or [BV,T,P,C], got {tuple(x.shape)}")

    def _publish_attention(self, cs: engine.CallState, BV: int, T: int, P: int):
        """attn_matrix of the LAST video-view, [T, E, P], pushed through attn_holder for forward hooks."""
        if self.fwb:
            return                           # no attention maps: the patch tokens are not used
        ca = self.pooling.cross_att
        if not ca.visual or cs.head_save is None:
            return
        attn = cs.plan.region(cs.head_save, "attn").view(-1, T, self.spec.n_entities, P)
        ca.attn_matrix = attn[-1].detach()
        ca.attn_holder(ca.attn_matrix)

    def make_call_state(self, project: bool = False, cls_emb=None) -> engine.CallState:
        running, tracked = self.bn_buffers()
        training = self.training
        seed = engine.new_seed() if (training and self.spec.drop_p > 0) else 0
        return engine.CallState(spec=self.spec, opts=self.run_options, training=training, bn_running=running,
                                bn_tracked=tracked, project=project, seed=seed,
                                seed_dev=self.seed_dev if (training and self.spec.drop_p > 0) else None,
                                cls_emb=cls_emb if self.fwb else None)

    def forward(self, x, video_masks=None, cls_emb=None):
        tokens = self.to_token_major(x.detach())
        BV, T, P, _ = tokens.shape
        cs = self.make_call_state(cls_emb=cls_emb)
        out = engine.HeadFn.apply(tokens, video_masks, cs, *self.head_params())
        self.last_call = cs
        self._publish_attention(cs, BV, T, P)
        return out
