"""Host logic of the packed / padding-free path against the values the
reference's own tests pin (tests/utils/test_packing.py:1489-1525, :1545-1572, :1575-1584) and the
metadata contract of unsloth/utils/packing.py:586-606."""
import torch

from unsloth_b200.packing import (clear_packed_caches, get_packed_info_from_kwargs,
                                  mask_packed_boundary_labels, mask_packed_sequence_boundaries,
                                  num_items_in_batch, packed_position_ids)


def test_boundary_mask_values_pinned_by_reference_tests():
    labels = torch.arange(8, dtype=torch.long).view(1, 8)
    out = mask_packed_boundary_labels(labels, torch.tensor([3, 5], dtype=torch.int32))
    # the shifted view the CE sees must be [1,2,-100,4,5,6,7] (test_packing.py:1522)
    assert out.reshape(-1).tolist() == [-100, 1, 2, -100, 4, 5, 6, 7]
    assert labels.reshape(-1).tolist() == list(range(8))          # out of place
    lengths = torch.tensor([2, 1, 3], dtype=torch.int32)
    once = mask_packed_boundary_labels(torch.arange(6).view(1, 6), lengths)
    assert once.reshape(-1).tolist() == [-100, 1, -100, -100, 4, 5]   # :1584
    assert torch.equal(mask_packed_boundary_labels(once, lengths), once)
    assert mask_packed_boundary_labels(labels, None) is labels
    assert mask_packed_boundary_labels(labels, torch.tensor([], dtype=torch.int32)) is labels


def test_boundary_mask_cases_of_the_reference_tests():
    """tests/utils/test_packing.py:1375-1436, case by case."""
    labels = torch.arange(6, dtype=torch.long).view(1, 6)
    out = mask_packed_boundary_labels(labels, torch.tensor([2, 1, 3], dtype=torch.int32))
    assert out.shape == labels.shape and out.dtype == labels.dtype
    # the two entry points mask exactly the same CE targets (:1388-1406)
    labels = torch.arange(100, 112, dtype=torch.long).view(1, 12)
    lengths = torch.tensor([5, 4, 3], dtype=torch.int32)
    shift_a = torch.empty_like(labels)
    shift_a[..., :-1] = labels[..., 1:]
    shift_a[..., -1] = -100
    assert mask_packed_sequence_boundaries(shift_a, lengths) is True
    masked = mask_packed_boundary_labels(labels, lengths)
    shift_b = torch.empty_like(masked)
    shift_b[..., :-1] = masked[..., 1:]
    shift_b[..., -1] = -100
    assert torch.equal(shift_a, shift_b)
    # idempotent on TRL's labels[position_ids == 0] = -100 (:1409-1420)
    lengths = torch.tensor([2, 1, 3], dtype=torch.int32)
    labels = torch.arange(6, dtype=torch.long).view(1, 6)
    trl = labels.clone()
    trl[torch.tensor([[0, 1, 0, 0, 1, 2]]) == 0] = -100
    once = mask_packed_boundary_labels(trl, lengths)
    assert torch.equal(once, trl) and torch.equal(mask_packed_boundary_labels(once, lengths), once)
    # no-op without packing (:1423-1427)
    assert mask_packed_boundary_labels(None, torch.tensor([2, 4])) is None
    assert mask_packed_sequence_boundaries(labels.clone(), None) is False
    # pad_to_multiple_of tail stays -100, no index out of range (:1430-1434)
    padded = torch.tensor([[10, 11, 12, 13, -100, -100]], dtype=torch.long)
    out = mask_packed_boundary_labels(padded, torch.tensor([2, 2], dtype=torch.int32))
    assert out.reshape(-1).tolist() == [10, 11, -100, 13, -100, -100]
    # lengths covering the whole row: the redirect must not corrupt a real target (:1437-1441)
    out = mask_packed_boundary_labels(torch.arange(4, dtype=torch.long).view(1, 4), [2, 2])
    assert out.reshape(-1).tolist() == [-100, 1, -100, 3]


def test_num_items_rule():
    # docs [10,11] [12] [13,14,15] -> 1 + 0 + 2 real CE targets (test_packing.py:1557-1572)
    labels = torch.tensor([[10, 11, 12, 13, 14, 15]])
    assert num_items_in_batch(labels, torch.tensor([2, 1, 3], dtype=torch.int32)) == 3
    assert num_items_in_batch(labels) == 5
    masked = mask_packed_boundary_labels(labels, torch.tensor([2, 1, 3]))
    assert int((masked[..., 1:] != -100).sum()) == 3               # the mask and the rule agree


def test_packed_info_and_position_ids():
    clear_packed_caches()
    lens = torch.tensor([3, 5], dtype=torch.int32)
    kw = {"packed_seq_lengths": lens}
    lengths, cu, mx = get_packed_info_from_kwargs(kw, "cpu")
    assert lengths.dtype == torch.int32 and cu.dtype == torch.int32
    assert cu.tolist() == [0, 3, 8] and mx == 5
    assert get_packed_info_from_kwargs(kw, "cpu")[1] is cu          # cached on tensor identity
    assert get_packed_info_from_kwargs({}, "cpu") is None
    assert packed_position_ids(lens).tolist() == [0, 1, 2, 0, 1, 2, 3, 4]
    # trailing pad tokens: one more segment so the boundaries cover the flattened row
    clear_packed_caches()
    _, cu2, mx2 = get_packed_info_from_kwargs(kw, "cpu", total=10)
    assert cu2.tolist() == [0, 3, 8, 10] and mx2 == 5
    assert packed_position_ids(lens, 10).tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 0, 1]
    try:
        clear_packed_caches()
        get_packed_info_from_kwargs(kw, "cpu", total=7)
        assert False, "over-long lengths must raise"
    except ValueError:
        pass
    clear_packed_caches()


def test_keep_dequant_policy_switch(monkeypatch):
    """UB200_KEEP_DEQUANT=0/1 forces the residency policy of the dequantised weights; without a
    CUDA device the automatic default is off (the policy keys on >= 128 GiB of device memory)."""
    from unsloth_b200.kernels import utils as KU
    try:
        for env, want in (("0", False), ("off", False), ("1", True), ("on", True)):
            monkeypatch.setenv("UB200_KEEP_DEQUANT", env)
            KU.set_keep_dequant(None)
            assert KU.keep_dequant() is want
        monkeypatch.delenv("UB200_KEEP_DEQUANT")
        KU.set_keep_dequant(None)
        if not torch.cuda.is_available():
            assert KU.keep_dequant() is False
        KU.set_keep_dequant(True)
        assert KU.keep_dequant() is True
    finally:
        KU.set_keep_dequant(None)
