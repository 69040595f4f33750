"""Sequence Contrastive Loss with the reference's `algos/` interface (CARL_MVF/algos/scl.py:18-105).

`compute_loss(model, videos, seq_lens, chosen_steps, video_masks, training)` and
`compute_sequence_loss(embs, seq_lens, steps, masks)` keep their signatures and return `{"loss": tensor}`;
the similarity / softmax / Gaussian-KL / gradient all run in csrc/scl.cu (engine.SCLFn), not in PyTorch.
"""
from __future__ import annotations

import torch

from .. import engine


class SCL(object):
    def __init__(self, cfg):
        self.cfg = cfg
        self.positive_type = cfg.SCL.POSITIVE_TYPE
        self.negative_type = cfg.SCL.NEGATIVE_TYPE
        self.temperature = cfg.SCL.SOFTMAX_TEMPERATURE
        self.label_varience = cfg.SCL.LABEL_VARIENCE
        self.embedding_size = cfg.MODEL.EMBEDDER_MODEL.EMBEDDING_SIZE
        self.positive_window = cfg.SCL.POSITIVE_WINDOW if "POSITIVE_WINDOW" in cfg.SCL else None
        if self.positive_type != "gauss":
            # the reference silently yields an all-zero label (loss 0) for anything else (scl.py:83-96)
            raise NotImplementedError("only SCL.POSITIVE_TYPE: gauss is implemented")
        if self.negative_type not in ("single_noself", "batch_noself"):
            raise NotImplementedError(f"SCL.NEGATIVE_TYPE {self.negative_type!r}: only single_noself / batch_noself "
                                      "(every configs_mvf/*.yml) are implemented")
        self.quirk = True   # keep weight 1e-6 on masked columns exactly like scl.py:80

    def compute_loss(self, model, videos, seq_lens, chosen_steps, video_masks=None, training=True):
        num_frames = self.cfg.TRAIN.NUM_FRAMES
        batch_size, num_views = videos.shape[0], videos.shape[1]
        num_steps = videos.shape[2]
        videos = videos.reshape((batch_size * num_views, num_steps) + tuple(videos.shape[3:]))
        if video_masks is not None:
            video_masks = video_masks.view(-1, 1, num_steps)
        embs = model(videos, num_frames, video_masks=video_masks, project=self.cfg.MODEL.PROJECTION)
        embs = embs.view(batch_size, num_views, num_frames, embs.size(-1))
        seq_lens = seq_lens.view(batch_size, num_views)
        return self.compute_sequence_loss(embs, seq_lens.to(embs.device), chosen_steps.to(embs.device),
                                          video_masks.to(embs.device))

    def compute_sequence_loss(self, embs, seq_lens, steps, masks=None):
        batch_size, num_views, num_frames, _ = embs.shape
        if masks is None:
            masks = torch.ones(batch_size * num_views, 1, num_frames, device=embs.device)
        loss = engine.SCLFn.apply(embs, seq_lens, steps, masks, self.temperature, self.label_varience,
                                  self.negative_type, self.quirk)
        return {"loss": loss}
