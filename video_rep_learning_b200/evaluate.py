"""Whole-video embedding extraction with the reference's chunking (CARL_MVF/evaluate.py:27-81, get_embeddings_dataset).

The reference feeds one video at a time, cut into ceil(seq_len / EVAL.FRAMES_PER_BATCH) equal chunks of up to 1000-2000
frames; each chunk goes through `model(curr_data, num_steps)` in eval mode -- no masks, no projection head, L2-normalised,
BatchNorm running statistics, positional encoding interpolated to the chunk length (models/utils.py:136-141).  With E
entities a chunk is a temporal sequence of S = E * frames tokens (6-12 k), which the fused head runs through the tcgen05
flash-attention kernels (csrc/attention_fa.cu); nothing of size S x S is ever materialised.

Only the hot-path call is mirrored here; the evaluation metrics (Kendall's tau, retrieval, classification on the
embeddings) stay CPU analytics outside the scope of this package.
"""
from __future__ import annotations

import math
from typing import List

import torch


def chunk_steps(seq_len: int, max_frames_per_batch: int, num_contexts: int = 1, context_stride: int = 1) -> List[torch.Tensor]:
    """The frame indices of every chunk, exactly as evaluate.py:45-55 builds them (INT, bit-exact)."""
    num_batches = int(math.ceil(float(seq_len) / max_frames_per_batch))
    frames_per_batch = int(math.ceil(float(seq_len) / num_batches))
    out = []
    for i in range(num_batches):
        curr_idx = i * frames_per_batch
        num_steps = min(seq_len - curr_idx, frames_per_batch)
        steps = torch.arange(curr_idx, curr_idx + num_steps)
        if num_contexts != 1:
            steps = steps.view(-1, 1) + context_stride * torch.arange(-(num_contexts - 1), 1).view(1, -1)
        out.append(torch.clamp(steps.view(-1), 0, seq_len - 1))
    return out


@torch.no_grad()
def embed_video(cfg, model, video: torch.Tensor) -> torch.Tensor:
    """video: frames [1, L, 3, H, W] or pre-computed patch tokens [1, L, P, C_in] (token-major) -> embeddings [L, D] on the
    host, chunk by chunk like get_embeddings_dataset.  The model must be in eval mode (BatchNorm running statistics)."""
    assert video.size(0) == 1, "evaluation feeds one video at a time (evaluate.py:41)"
    if model.training:
        raise RuntimeError("embed_video: call model.eval() first (evaluate.py:38)")
    seq_len = int(video.size(1))
    ctx = int(cfg.DATA.NUM_CONTEXTS) if "NUM_CONTEXTS" in cfg.DATA else 1
    stride = int(cfg.DATA.CONTEXT_STRIDE) if "CONTEXT_STRIDE" in cfg.DATA else 1
    embs = []
    for steps in chunk_steps(seq_len, int(cfg.EVAL.FRAMES_PER_BATCH), ctx, stride):
        curr = video[:, steps.to(video.device)]
        num_steps = int(steps.numel()) // ctx
        feats = model(curr, num_steps)
        embs.append(feats[0].float().cpu())
    return torch.cat(embs, dim=0)


def get_embeddings_dataset(cfg, model, data_loader) -> dict:
    """Same name, arguments and return value as evaluate.py:27-81: one pass over a batch-size-1 loader of
    (video, frame_label, seq_len, chosen_steps, video_masks, names); frames whose label is negative are dropped.
    Returns {'embs', 'labels', 'seq_lens', 'input_lens', 'steps', 'names'} with numpy arrays per video."""
    embs_list, labels_list, seq_lens_list, input_lens_list, steps_list, names_list = [], [], [], [], [], []
    model.eval()
    for video, frame_label, seq_len, chosen_steps, _video_masks, names in data_loader:
        assert video.size(0) == 1                                             # evaluate.py:41
        assert video.size(1) == frame_label.size(1) == int(seq_len.item())    # evaluate.py:42
        embs = embed_video(cfg, model, video)
        valid = frame_label[0] >= 0
        embs_list.append(embs[valid.cpu()].numpy())
        labels_list.append(frame_label[0][valid].cpu().numpy())
        seq_lens_list.append(int(seq_len.item()))
        input_lens_list.append(len(video[0]))
        steps_list.append(chosen_steps[0].cpu().numpy())
        names_list.append(names[0])
    return {"embs": embs_list, "labels": labels_list, "seq_lens": seq_lens_list, "input_lens": input_lens_list,
            "steps": steps_list, "names": names_list}
