"""Minimal cfg object compatible with how the reference reads its EasyDict cfg on this path
(attribute access, item access, `'KEY' in cfg.NODE`; CARL_MVF/utils/config.py, utils/parser.py:64-103).

The reference's config *system* (defaults, --opts overlay, per-logdir pinning) is out of scope; this is only
the container the hot-path modules need, plus a YAML loader so a configs_mvf/*.yml from the reference tree can
be used as is.
"""
from __future__ import annotations

import copy

import yaml


class Cfg(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Cfg):
            v = Cfg(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


def load_yaml(path: str) -> Cfg:
    with open(path) as f:
        return Cfg(yaml.safe_load(f))


def mvf_cfg(*, c_in=2304, network="TIMM-vit_base_patch16_224.dino", smart_feats="3,7,11", num_frames=20, batch_size=1,
            entities=3, capacity=2, emb=128, final="one", one_hot="pool", drop=0.1, hidden=256, d_ff=1024, heads=8,
            layers=3, fc_layers=((256, True), (256, True)), projection_size=128, negative_type="single_noself",
            pool_channels=None) -> Cfg:
    """The hot-path subset of configs_mvf/penn_mvf.yml (defaults) / fg99_mvf.yml (entities=6, capacity=6, emb=256,
    final='avg', smart_feats='9,10,11'), with the run-time key MODEL.BASE_MODEL.OUT_CHANNEL filled in."""
    em = dict(HIDDEN_SIZE=hidden, D_FF=d_ff, NUM_HEADS=heads, NUM_LAYERS=layers, CAPACITY_SCALAR=capacity,
              EMBEDDING_SIZE=emb, FC_DROPOUT_RATE=drop, FC_LAYERS=[list(x) for x in fc_layers], FUSION_TYPE="smart",
              SMART_TOKENS=entities, SMART_ONE_HOT=one_hot, SMART_FEATS=smart_feats, SMART_FINAL=final)
    if pool_channels is not None:
        em["SMART_POOL_CHANNELS"] = pool_channels
    return Cfg(dict(
        TRAINING_ALGO="scl", USE_AMP=True, RNG_SEED=1, LOGDIR="/tmp/mvf_b200_logs",
        MODEL=dict(BASE_MODEL=dict(NETWORK=network, LAYER=12, FRAMES_PER_BATCH=40, OUT_CHANNEL=c_in), EMBEDDER_MODEL=em,
                   EMBEDDER_TYPE="transformer", L2_NORMALIZE=True, PROJECTION=True, PROJECTION_HIDDEN_SIZE=512,
                   PROJECTION_SIZE=projection_size, TRAIN_BASE="frozen"),
        SCL=dict(LABEL_VARIENCE=10.0, POSITIVE_TYPE="gauss", NEGATIVE_TYPE=negative_type, SOFTMAX_TEMPERATURE=0.1,
                 POSITIVE_WINDOW=5),
        TRAIN=dict(BATCH_SIZE=batch_size, MAX_EPOCHS=1, NUM_FRAMES=num_frames),
        DATA=dict(SAMPLING_STRATEGY="time_augment", SAMPLING_REGION=1.5, CONSISTENT_OFFSET=0.2, NUM_CONTEXTS=1,
                  CONTEXT_STRIDE=1),
        OPTIMIZER=dict(GRAD_CLIP=10, TYPE="AdamOptimizer", WEIGHT_DECAY=1e-5, LR=dict(INITIAL_LR=1e-4)),
    ))
