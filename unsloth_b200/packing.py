"""Packed-sequence label masking handed to the fused CE (host side).

Semantics of the reference's `mask_packed_boundary_labels`
(unsloth/utils/packing.py:733-772), pinned by its own test
tests/utils/test_packing.py:1489-1525: on RAW (unshifted) labels, out of place, every first
token of a following document (`labels[cumsum(lengths)]`) becomes ignore_index; the final
cumsum (== total) is redirected to slot 0, which the internal shift discards.
"""
from __future__ import annotations

import torch


def mask_packed_boundary_labels(labels, seq_lengths, *, ignore_index: int = -100):
    if labels is None or not isinstance(labels, torch.Tensor) or seq_lengths is None:
        return labels
    lengths = torch.as_tensor(seq_lengths, device=labels.device).to(torch.int64).reshape(-1)
    total = labels.numel()
    if lengths.numel() == 0 or total == 0:
        return labels
    starts = torch.cumsum(lengths, dim=0)
    starts = torch.where(starts < total, starts, torch.zeros_like(starts))
    return labels.reshape(-1).index_fill(0, starts, ignore_index).view(labels.shape)
