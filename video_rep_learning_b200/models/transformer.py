"""`TransformerModel` -- backbone glue + MV-Former head + projection, with the reference's interface
(CARL_MVF/models/transformer.py:16-244).

Scope: the frozen ViT backbone is an upstream PyTorch producer and is NOT part
of the product; this class keeps the hand-off contract (`forward(x, num_frames, video_masks, project,
classification)`, attributes `.backbone .res_finetune .embed .ssl_projection .embedding_size`) and replaces
what happens after the backbone:
  * tokens stay token-major [n, P, C] as the ViT emits them -- the reference's movedim/reshape/.contiguous()
    into NCHW (transformer.py:203-213) and the permute straight back (mvformer.py:244-246) are gone;
  * head + MLPHead + F.normalize run as ONE autograd node (engine.ModelFn): one flat gradient buffer, one
    NCCL all-reduce, BatchNorm statistics exchanged across ranks when torch.distributed is initialised.
Only FUSION_TYPE: smart with a frozen backbone (TRAIN_BASE: frozen, every configs_mvf/*.yml) is supported.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import engine
from .mvformer import MultiEntityTransformerEmbModel, _get
from .resnet_c2d import MLPHead

_TIMM_WIDTH = {
    "vit_small_patch16_224.dino": 384, "vit_small_patch8_224.dino": 384, "vit_small_patch14_dinov2.lvd142m": 384,
    "vit_base_patch16_224.dino": 768, "vit_base_patch8_224.dino": 768, "vit_base_patch14_dinov2.lvd142m": 768,
    "vit_large_patch14_dinov2.lvd142m": 1024, "vit_giant_patch14_dinov2.lvd142m": 1536,
}


class FeatureExtractor(nn.Module):
    """Collects the token outputs of selected ViT blocks (role of transformer.py:306-333).

    The reference concatenates the k hooked block outputs along channels (`torch.cat`, one full write + read of the
    largest tensor on the path) and `TransformerModel.forward` then copies the result again while dropping CLS and
    re-laying it out.  Here the hooks write STRAIGHT into the head's token buffer: `bind(dest)` hands over the destination
    slice [n, P, C*k] (token-major, CLS dropped, any dtype -- the bf16 conversion rides on the same copy) and every hook
    stores its block's patch tokens into its own channel range of it as the block finishes.  Nothing else is materialised:
    no concatenated tensor, no second copy.  Without a bound destination `forward` returns (tokens [n, 1+P, C*k], cls) like
    the reference (one `torch.cat`), which is what third-party callers expect."""

    def __init__(self, vit: nn.Module, layers):
        super().__init__()
        self.vit = vit
        self.layers = [int(l) for l in layers]
        self._feats = {}
        self._dest = None
        for i, l in enumerate(self.layers):
            vit.blocks[l].register_forward_hook(self._make_hook(i, l))

    def bind(self, dest):
        """dest: [n, P, C*k] view the next forward writes the patch tokens into (None: unbind)."""
        self._dest = dest

    def _make_hook(self, i, l):
        def hook(_m, _i, out):
            if self._dest is not None:
                C = out.shape[-1]
                self._dest[:, :, i * C:(i + 1) * C].copy_(out[:, 1:, :])     # drop CLS, cast, place: one strided copy
            else:
                self._feats[l] = out
        return hook

    def forward(self, x):
        self._feats.clear()
        final = self.vit.forward_features(x)
        if self._dest is not None:
            return None, final[:, 0]
        toks = torch.cat([self._feats[l] for l in self.layers], dim=-1)
        return toks, final[:, 0]


class TransformerModel(nn.Module):
    def __init__(self, cfg, local_rank=None, backbone: Optional[nn.Module] = None):
        super().__init__()
        self.cfg = cfg
        em = cfg.MODEL.EMBEDDER_MODEL
        self.fusion_type = _get(em, "FUSION_TYPE", "late")
        if self.fusion_type != "smart":
            raise NotImplementedError("only MODEL.EMBEDDER_MODEL.FUSION_TYPE: smart (MV-Former) is built here; the "
                                      "late-fusion CARL head is out of scope (SURVEY.md section 2, row 5)")
        if _get(cfg.MODEL, "CLS_RES", False):
            raise NotImplementedError("MODEL.CLS_RES is an optional branch outside the hot path (SURVEY.md section 8f)")
        if cfg.MODEL.TRAIN_BASE != "frozen":
            raise NotImplementedError("only TRAIN_BASE: frozen is supported: tokens are consumed without a gradient path")
        self.use_cls_res = False
        self.backbone_type = "timm"
        net = str(cfg.MODEL.BASE_MODEL.NETWORK)
        feats = [int(s) for s in str(_get(em, "SMART_FEATS", "11")).split(",")]
        if backbone is not None:
            self.backbone = backbone          # any module: frames [n,3,H,W] -> (tokens [n,1+P,C_in], cls [n,C])
            # the shipped yml files carry no OUT_CHANNEL: the reference derives it from NETWORK (transformer.py:40-56, 90,
            # 119-133), so an unmodified config works with a caller-supplied producer as well
            if "OUT_CHANNEL" not in cfg.MODEL.BASE_MODEL:
                if net.startswith("TIMM-") and net[5:] in _TIMM_WIDTH:
                    cfg.MODEL.BASE_MODEL.OUT_CHANNEL = _TIMM_WIDTH[net[5:]] * len(feats)
                elif "resnet" in net.lower():
                    cfg.MODEL.BASE_MODEL.OUT_CHANNEL = 2048
                else:
                    raise ValueError(f"MODEL.BASE_MODEL.OUT_CHANNEL is not set and cannot be derived from NETWORK: {net}")
        elif net.startswith("TIMM-"):
            name = net[5:]
            if name not in _TIMM_WIDTH:
                raise ValueError(f"unknown/unsupported TIMM model: {name}")
            try:
                import timm
            except ImportError as e:  # pragma: no cover - timm is not in the build image
                raise RuntimeError("timm is not installed: pass `backbone=` (frames -> (tokens, cls)) explicitly") from e
            vit = timm.create_model(name, pretrained=True)
            cfg.MODEL.BASE_MODEL.OUT_CHANNEL = _TIMM_WIDTH[name] * len(feats)
            self.backbone = FeatureExtractor(vit, feats)
        else:
            raise NotImplementedError("ResNet backbones belong to the original CARL models (out of scope)")
        for p in self.backbone.parameters():
            p.requires_grad = False
        self.res_finetune = nn.Identity()
        self.embed = MultiEntityTransformerEmbModel(cfg)
        self.embedding_size = self.embed.embedding_size
        if cfg.MODEL.PROJECTION:
            self.ssl_projection = MLPHead(cfg, spec=self.embed.spec)
        if cfg.TRAINING_ALGO == "classification":
            raise NotImplementedError("TRAINING_ALGO: classification is out of scope (SURVEY.md section 2, row 8)")

    # ---- options shared by the fused chain ---------------------------------------------------------------
    @property
    def run_options(self) -> engine.RunOptions:
        return self.embed.run_options

    def backbone_tokens(self, x: torch.Tensor, with_cls: bool = False, out: Optional[torch.Tensor] = None,
                        dtype: Optional[torch.dtype] = None):
        """frames [BV,T,3,H,W] -> patch tokens [BV,T,P,C_in] (token-major, CLS dropped), chunked over frames like
        transformer.py:175-189 (FRAMES_PER_BATCH), backbone in eval mode under no_grad.
        `out` (optional): the token buffer to fill -- e.g. graph.GraphedTrainStep.tokens, so the producer writes where the
        captured step reads; `dtype`: element type of a freshly allocated buffer (default: the backbone's).  With the
        package's FeatureExtractor the hooked ViT blocks write their patch tokens directly into that buffer (no concatenated
        tensor, no second copy); any other backbone goes through one slice copy.
        with_cls: also the CLS embeddings [BV*T, C] concatenated chunk by chunk exactly as transformer.py:200,217 does
        (chunk-major rows -- with T > FRAMES_PER_BATCH that differs from the video-major order the head assumes; the
        reference's own behaviour is kept)."""
        BV, T, c, h, w = x.shape
        fpb = self.cfg.MODEL.BASE_MODEL.FRAMES_PER_BATCH
        cls_chunks = []
        self.backbone.eval()
        direct = isinstance(self.backbone, FeatureExtractor) and isinstance(self.res_finetune, nn.Identity)
        if direct and out is None:
            n_patches = getattr(getattr(self.backbone.vit, "patch_embed", None), "num_patches", None)
            if n_patches is not None:
                par = next(self.backbone.vit.parameters(), None)
                out = torch.empty(BV, T, int(n_patches), int(self.cfg.MODEL.BASE_MODEL.OUT_CHANNEL),
                                  dtype=dtype or (par.dtype if par is not None else x.dtype), device=x.device)
        for i in range(int(math.ceil(float(T) / fpb))):
            t0 = i * fpb
            t1 = min(T, t0 + fpb)
            cur = x[:, t0:t1].contiguous().view(-1, c, h, w)
            if direct and out is not None and (BV == 1 or t1 - t0 == T):
                # the frames of this chunk are one contiguous [n, P, C_in] block of the buffer: hooks write in place
                self.backbone.bind(out[:, t0:t1].reshape(BV * (t1 - t0), out.shape[2], out.shape[3]))
                try:
                    with torch.no_grad():
                        _, _cls = self.backbone(cur)
                finally:
                    self.backbone.bind(None)
            else:
                with torch.no_grad():
                    toks, _cls = self.backbone(cur)
                toks = self.res_finetune(toks)
                n, ntok, C = toks.shape
                if out is None:
                    out = torch.empty(BV, T, ntok - 1, C, dtype=dtype or toks.dtype, device=toks.device)
                out[:, t0:t1].copy_(toks[:, 1:, :].reshape(BV, t1 - t0, ntok - 1, C))   # drop CLS, keep token-major
            if with_cls:
                cls_chunks.append(_cls)
        if with_cls:
            return out, torch.cat(cls_chunks, dim=0)
        return out

    def forward(self, x, num_frames=None, video_masks=None, project=False, classification=False):
        if classification:
            raise NotImplementedError("classification head is out of scope")
        cls_emb = None
        if x.dim() == 5 and x.shape[2] == 3:
            if self.embed.fwb:
                tokens, cls_emb = self.backbone_tokens(x, with_cls=True)
            else:
                tokens = self.backbone_tokens(x)
        else:
            tokens = self.embed.to_token_major(x)
        return self.forward_tokens(tokens, video_masks=video_masks, project=project, cls_emb=cls_emb)

    def forward_tokens(self, tokens, video_masks=None, project=False, cls_emb=None):
        """The hot path proper: patch tokens [BV,T,P,C_in] -> embeddings [BV,T,D] (transformer.py:219-230)."""
        cfg = self.cfg
        BV, T, P, _ = tokens.shape
        use_proj = bool(cfg.MODEL.PROJECTION and project)
        if not use_proj and not cfg.MODEL.L2_NORMALIZE:
            return self.embed(tokens, video_masks=video_masks, cls_emb=cls_emb)
        cs = self.embed.make_call_state(project=1 if use_proj else 0, cls_emb=cls_emb)
        params = list(self.embed.head_params())
        if cfg.MODEL.PROJECTION:
            params += self.ssl_projection.proj_params()
            pr, pt = self.ssl_projection.bn_buffers()
            cs.bn_running = cs.bn_running + pr
            cs.bn_tracked = cs.bn_tracked + pt
        else:
            params += [None] * 6
            cs.bn_running = cs.bn_running + [None, None]
            cs.bn_tracked = cs.bn_tracked + [None]
        out = engine.ModelFn.apply(tokens.detach(), video_masks, cs, *params)
        self.embed.last_call = cs
        self.embed._publish_attention(cs, BV, T, P)
        return out
