"""Packed / padding-free batches (host side): `packed_seq_lengths` ->
(lengths, cu_seqlens, max_seqlen) for the varlen attention call, reset-style position ids for the
indexed RoPE kernel, boundary label masking handed to the fused CE, and the item count rule.

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


def mask_packed_sequence_boundaries(shift_labels, seq_lengths, *, ignore_index: int = -100) -> bool:
    """The in-place guard on ALREADY SHIFTED labels (unsloth/utils/packing.py:710-730), used by the
    logits path (`fast_cross_entropy_loss` after the caller's shift, llama.py:1545-1552): the last
    token of every packed document must not predict the next document's first token.  Returns
    whether anything was masked."""
    if seq_lengths is None:
        return False
    lengths = torch.as_tensor(seq_lengths, device=shift_labels.device).to(torch.int64).reshape(-1)
    if lengths.numel() == 0:
        return False
    flat = shift_labels.reshape(-1)
    boundary = torch.cumsum(lengths, dim=0) - 1
    boundary = boundary[boundary < flat.shape[0]]
    if boundary.numel() == 0:
        return False
    flat[boundary] = ignore_index
    return True


_PACKED_INFO_CACHE = {}


def get_packed_info_from_kwargs(kwargs, device, total=None):
    """unsloth/utils/packing.py:586-606: (lengths int32[n], cu_seqlens int32[n+1], max_seqlen) on
    `device`, cached on the identity of the `packed_seq_lengths` tensor so that the 32 layers of a
    step share one result.  The collator hands the lengths over on the CPU
    (packing.py:232, :277), so max_seqlen is read there: no device->host sync in the step.
    With `total` (the flattened token count) given, pad tokens after the last document become one
    more segment (as packing.py:403-427 does), so every token has a defined attention output."""
    seq_lengths = kwargs.get("packed_seq_lengths")
    if seq_lengths is None:
        return None
    dev = torch.device(device)
    entry = _PACKED_INFO_CACHE.get(dev)
    if entry is not None and entry["seq_lengths"] is seq_lengths:
        return entry["result"]
    src = torch.as_tensor(seq_lengths).reshape(-1)
    if src.numel() == 0:
        return None
    if total is not None:
        covered = int(src.sum().item())
        if covered > int(total):
            raise ValueError("packed_seq_lengths sum to %d > %d tokens in the batch" % (covered, total))
        if covered < int(total):
            src = torch.cat([src.cpu(), torch.tensor([int(total) - covered], dtype=src.dtype)])
    max_seqlen = int(src.max().item())              # CPU tensor in the collator route
    lengths = src.to(device=dev, dtype=torch.int32, non_blocking=True)
    cu_seqlens = torch.zeros(lengths.numel() + 1, dtype=torch.int32, device=dev)
    torch.cumsum(lengths, dim=0, dtype=torch.int32, out=cu_seqlens[1:])
    result = (lengths, cu_seqlens, max_seqlen)
    _PACKED_INFO_CACHE[dev] = {"seq_lengths": seq_lengths, "result": result}
    return result


def clear_packed_caches():
    """packing.py:775-780."""
    _PACKED_INFO_CACHE.clear()


def packed_position_ids(seq_lengths, total=None, device=None):
    """Reset-style position ids of a flattened packed row (what a padding-free collator with
    `return_position_ids` emits): 0..len_0-1, 0..len_1-1, ...; when sum(lengths) < total the
    trailing pad tokens form one more segment, like packing.py:403-427.  int32 [total]."""
    lengths = torch.as_tensor(seq_lengths).reshape(-1).to(torch.int64).cpu()
    covered = int(lengths.sum())
    total = covered if total is None else int(total)
    if covered < total:
        lengths = torch.cat([lengths, torch.tensor([total - covered])])
    starts = torch.cumsum(lengths, 0) - lengths
    pos = torch.arange(int(lengths.sum())) - torch.repeat_interleave(starts, lengths)
    pos = pos[:total].to(torch.int32)
    return pos.to(device) if device is not None else pos


def num_items_in_batch(labels, seq_lengths=None, ignore_index: int = -100):
    """The count unsloth_zoo derives for token-mean loss (tests/utils/test_packing.py:1545-1552):
    shifted non-ignored targets, minus the N-1 boundary targets of a packed row."""
    count = int((labels[..., 1:] != ignore_index).sum())
    if seq_lengths is not None:
        count -= int(torch.count_nonzero(torch.as_tensor(seq_lengths) > 0)) - 1
    return count
