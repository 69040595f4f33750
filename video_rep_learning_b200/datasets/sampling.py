"""Two-view temporal frame sampling (integer-exact host logic).

Restates `sample_frames` of CARL_MVF/datasets/{penn_action.py:152-206, finegym.py:167-221,
kinetics400.py:135-182, pouring.py:130-189}.  The draws come from the global numpy and torch generators in the
reference's order (np.random.uniform -> np.random.randint -> torch.randperm), so identical seeds give
bit-identical `steps / chosen_steps / video_mask`.  It stays on the host: it runs in DataLoader workers, costs
microseconds, and its outputs (int64 steps, float mask) are inputs of the SCL kernel.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

_BLOCK_RULE = ("penn_action", "kinetics400", "pouring", "finegym", "pouring_fix")


def sample_frames(seq_len: int, num_frames: int, pre_steps=None, *, dataset: str = "penn_action",
                  sampling_strategy: str = "time_augment", sampling_region: float = 1.5,
                  consistent_offset: float = 0.2, num_contexts: int = 1, context_stride: int = 1):
    """Returns (steps, chosen_steps, video_mask).

    dataset selects the sampling-block rule: ceil(r*seq_len) for penn_action / kinetics400 / pouring,
    ceil(r*num_valid) for finegym, ceil(r*num_frames) for pouring with SAMPLE_FIX ('pouring_fix').
    """
    if dataset not in _BLOCK_RULE:
        raise ValueError(f"unknown dataset {dataset!r}")
    pre_offset = min(pre_steps) if pre_steps is not None else None
    if sampling_strategy == "offset_uniform":
        if seq_len >= num_frames:
            steps = torch.sort(torch.randperm(seq_len)[:num_frames])[0]
        else:
            steps = torch.arange(0, num_frames)
    elif sampling_strategy == "time_augment":
        num_valid = min(seq_len, num_frames)
        ratio = np.random.uniform(low=1.0, high=sampling_region) if sampling_region > 1 else 1.0
        basis = num_valid if dataset == "finegym" else (num_frames if dataset == "pouring_fix" else seq_len)
        block_size = math.ceil(ratio * basis)
        if pre_steps is not None and consistent_offset != 0:
            shift = int((1 - consistent_offset) * num_valid)
            lo = max(0, min(seq_len - block_size, pre_offset - shift))
            hi = max(1, min(seq_len - block_size + 1, pre_offset + shift + 1))
            offset = np.random.randint(low=lo, high=hi)
        else:
            offset = np.random.randint(low=0, high=max(seq_len - block_size, 1))
        steps = torch.sort(offset + torch.randperm(block_size)[:num_valid])[0]
        if num_valid < num_frames:
            steps = F.pad(steps, (0, num_frames - num_valid), "constant", seq_len)
    else:
        raise ValueError("Sampling strategy %s is unknown. Supported values are stride, offset_uniform ." % sampling_strategy)
    video_mask = torch.ones(num_frames)
    video_mask[steps < 0] = 0
    video_mask[steps >= seq_len] = 0
    chosen_steps = torch.clamp(steps.clone(), 0, seq_len - 1)
    if num_contexts == 1:
        steps = chosen_steps
    else:
        steps = steps.view(-1, 1) + context_stride * torch.arange(-(num_contexts - 1), 1).view(1, -1)
        steps = torch.clamp(steps.view(-1), 0, seq_len - 1)
    return steps, chosen_steps, video_mask
