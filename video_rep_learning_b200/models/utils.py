"""Parameter containers of the temporal encoder, with the reference's module / parameter names.

Mirrors CARL_MVF/models/utils.py (module tree, constructor signatures, initialisation order) so that
`state_dict()` keys, shapes and seeded initial values are interchangeable with the reference
(SURVEY.md section 8b).  The arithmetic of these modules is NOT executed here: inside the MV-Former head
the whole encoder runs in libmvf_b200.so (engine.HeadFn / engine.ModelFn).  Calling a sub-module's forward
on its own raises, because a PyTorch re-implementation would be a silent non-CUDA-kernel path.
"""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn


class _FusedOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container: it is executed inside the fused MV-Former head "
            "(video_rep_learning_b200.models.mvformer.MultiEntityTransformerEmbModel), not on its own")


class MultiheadedAttention(_FusedOnly):
    """models/utils.py:47-108 -- linear_Q2d / linear_K2d / linear_V2d / linear_d2Q."""

    def __init__(self, d_model_Q, d_model_K, d_model_V, H, dout_p=0.0, d_model=None, d_out=None):
        super().__init__()
        self.d_model_Q, self.d_model_K, self.d_model_V, self.H = d_model_Q, d_model_K, d_model_V, H
        self.d_model = d_model if d_model is not None else d_model_Q
        self.d_out = d_out if d_out is not None else d_model_Q
        self.dout_p = dout_p
        assert self.d_model % H == 0
        self.d_k = self.d_model // H
        self.linear_Q2d = nn.Linear(d_model_Q, self.d_model)
        self.linear_K2d = nn.Linear(d_model_K, self.d_model)
        self.linear_V2d = nn.Linear(d_model_V, self.d_model)
        self.linear_d2Q = nn.Linear(self.d_model, self.d_out)
        self.dropout = nn.Dropout(dout_p)
        self.visual = False


class ResidualConnection(_FusedOnly):
    """models/utils.py:147-159 -- pre-LN residual branch; owns `norm`."""

    def __init__(self, size, dout_p):
        super().__init__()
        self.norm = nn.LayerNorm(size)
        self.dropout = nn.Dropout(dout_p)


class PositionwiseFeedForward(_FusedOnly):
    """models/utils.py:176-194 -- fc1 / ReLU / fc2."""

    def __init__(self, d_model, d_ff, dout_p):
        super().__init__()
        self.d_model, self.d_ff, self.dout_p = d_model, d_ff, dout_p
        self.fc1 = nn.Linear(d_model, d_ff)
        self.fc2 = nn.Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dout_p)
        self.activation = nn.ReLU(inplace=True)


class EncoderLayer(_FusedOnly):
    """models/utils.py:196-226; every matrix is re-initialised with xavier_uniform (utils.py:206-208)."""

    def __init__(self, d_model, dout_p, H=8, d_ff=None, d_hidden=None):
        super().__init__()
        self.res_layer0 = ResidualConnection(d_model, dout_p)
        self.res_layer1 = ResidualConnection(d_model, dout_p)
        d_hidden = d_model if d_hidden is None else d_hidden
        d_ff = 4 * d_model if d_ff is None else d_ff
        self.self_att = MultiheadedAttention(d_model, d_model, d_model, H, d_model=d_hidden)
        self.feed_forward = PositionwiseFeedForward(d_model, d_ff, dout_p=0.0)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class Encoder(_FusedOnly):
    """models/utils.py:228-242 -- N deep copies of ONE initialised layer (all layers start identical)."""

    def __init__(self, d_model, dout_p, H, d_ff, N, d_hidden=None):
        super().__init__()
        proto = EncoderLayer(d_model, dout_p, H, d_ff, d_hidden)
        self.enc_layers = nn.ModuleList([deepcopy(proto) for _ in range(N)])


class PositionalEncoder(_FusedOnly):
    """models/utils.py:128-145 -- parameter-free; the sin/cos table is generated on the device
    (csrc/elementwise.cu posenc_table_kernel) instead of a numpy loop + H2D copy per forward."""

    def __init__(self, cfg, d_model, dout_p, seq_len=3660):
        super().__init__()
        self.cfg = cfg
        self.d_model = d_model
        self.dropout = nn.Dropout(dout_p)
        self.seq_len = seq_len
