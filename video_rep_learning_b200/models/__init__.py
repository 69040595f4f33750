"""`models/` of the reference (CARL_MVF/models/__init__.py:8-59): build_model + checkpoint helpers."""
from __future__ import annotations

import os

import torch

from .mvformer import LearnableTokenPooling, LSTPCrossAtt, MultiEntityTransformerEmbModel, head_spec_from_cfg
from .resnet_c2d import MLPHead
from .transformer import TransformerModel
from .utils import Encoder, EncoderLayer, MultiheadedAttention, PositionalEncoder, PositionwiseFeedForward, ResidualConnection


def build_model(cfg, local_rank=None, backbone=None):
    """models/__init__.py:8-15; only the transformer embedder (MV-Former) exists in this package."""
    if cfg.MODEL.EMBEDDER_TYPE != "transformer":
        raise NotImplementedError("only MODEL.EMBEDDER_TYPE: transformer is built here")
    return TransformerModel(cfg, local_rank, backbone=backbone)


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def save_checkpoint(cfg, model, optimizer, epoch):
    """Same file layout as models/__init__.py:17-29 ({epoch, model_state, optimizer_state, cfg})."""
    path = os.path.join(cfg.LOGDIR, "checkpoints")
    os.makedirs(path, exist_ok=True)
    ckpt = {"epoch": epoch, "model_state": _unwrap(model).state_dict(), "optimizer_state": optimizer.state_dict(), "cfg": cfg}
    torch.save(ckpt, os.path.join(path, "checkpoint_epoch_{:05d}.pth".format(epoch)))


def load_checkpoint(cfg, model, optimizer):
    """Latest checkpoint in LOGDIR/checkpoints, else MODEL.PRETRAINED_CHECKPOINT (models/__init__.py:35-59)."""
    path = os.path.join(cfg.LOGDIR, "checkpoints")
    names = sorted(f for f in os.listdir(path) if "checkpoint" in f) if os.path.exists(path) else []
    if names:
        ckpt = torch.load(os.path.join(path, names[-1]), map_location="cpu", weights_only=False)
        _unwrap(model).load_state_dict(ckpt["model_state"])
        optimizer.load_state_dict(ckpt["optimizer_state"])
        return ckpt["epoch"] + 1
    pre = cfg.MODEL.PRETRAINED_CHECKPOINT if "PRETRAINED_CHECKPOINT" in cfg.MODEL else None
    if pre:
        ckpt = torch.load(pre, map_location="cpu", weights_only=False)
        _unwrap(model).load_state_dict(ckpt["model_state"], strict=False)
    return 0
