"""SimCLR-style projection head `MLPHead` with the reference's API (CARL_MVF/models/resnet_c2d.py:112-126).

Only the class on the MV-Former hot path is provided; the ResNet50 / conv embedder baselines of the original
CARL (resnet_c2d.py:1-110, 128-236) are out of scope.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from .. import engine


class MLPHead(nn.Module):
    """Linear(D, PROJECTION_SIZE) -> BatchNorm1d -> ReLU -> Linear(PROJECTION_SIZE, D).

    NB: like the reference, the hidden width is MODEL.PROJECTION_SIZE; PROJECTION_HIDDEN_SIZE is unused.
    forward(x [b, l, D]) returns the un-normalised projection, as the reference does; TransformerModel fuses the
    following F.normalize into the same kernel chain (engine.ModelFn).
    """

    def __init__(self, cfg, spec: engine.HeadSpec = None):
        super().__init__()
        projection_hidden_size = cfg.MODEL.PROJECTION_SIZE
        self.embedding_size = cfg.MODEL.EMBEDDER_MODEL.EMBEDDING_SIZE
        self.net = nn.Sequential(nn.Linear(self.embedding_size, projection_hidden_size),
                                 nn.BatchNorm1d(projection_hidden_size),
                                 nn.ReLU(True),
                                 nn.Linear(projection_hidden_size, self.embedding_size))
        if spec is None:
            from .mvformer import head_spec_from_cfg
            spec = head_spec_from_cfg(cfg)
        self.spec = spec
        self.run_options = engine.RunOptions()

    def proj_params(self) -> List[torch.Tensor]:
        n = self.net
        return [n[0].weight, n[0].bias, n[1].weight, n[1].bias, n[3].weight, n[3].bias]

    def bn_buffers(self):
        bn = self.net[1]
        return [bn.running_mean, bn.running_var], [bn.num_batches_tracked]

    def forward(self, x):
        # standalone use, as the reference calls it (transformer.py:227): projection without the normalisation
        n_fc = len(self.spec.fc_channels)
        running, tracked = self.bn_buffers()
        cs = engine.CallState(spec=self.spec, opts=self.run_options, training=self.training,
                              bn_running=[None] * (2 * n_fc) + running, bn_tracked=[None] * n_fc + tracked, project=2)
        return engine.ProjFn.apply(x, cs, *self.proj_params())
