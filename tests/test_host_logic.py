"""CPU tier: host-side logic of the data path that needs no kernel -- batch layout, interpolation coefficients, triplet
gathering -- against the oracle restatements (themselves pinned against the reference in test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import aesr_oracle as O


def test_prepare_batch_pairs_layout_matches_oracle():
    """datasets/common_brains.py:285-321: [B,3,H,W] -> image [2B,1,H,W] (all "from", then all "to"), slice_between."""
    from superresolution_aniso_mri_b200 import evaluation as E
    b = torch.rand(4, 3, 8, 8, generator=torch.Generator().manual_seed(1))
    want = O.prepare_batch_pairs(b)
    got = E.prepare_batch_pairs({"image": b.clone()})
    assert torch.equal(got["image"], want["image"]) and torch.equal(got["slice_between"], want["slice_between"])
    two = E.prepare_batch_pairs({"image": b[:, :2].clone()})
    assert two["image"].shape == (8, 1, 8, 8) and "slice_between" not in two
    sp = E.prepare_batch_pairs({"image": b.clone()}, expand_type="split")
    assert torch.equal(sp["image_from"], b[:, 0:1]) and torch.equal(sp["image_to"], b[:, 1:2]) and sp["image"].shape == b.shape
    with pytest.raises(ValueError):
        E.prepare_batch_pairs({"image": b}, expand_type="reshape")
    with pytest.raises(AssertionError):
        E.prepare_batch_pairs({"image": b[:3]})                       # odd batch, like the reference's assert


def test_interpolation_coefficients_match_oracle():
    from superresolution_aniso_mri_b200 import evaluation as E
    f, t, m = np.array([3, 10, 8, 0]), np.array([7, 6, 12, 4]), np.array([4, 8, 11, 1])
    a1, a2 = E.determine_interpol_coefficients(f, t, m)
    o1, o2 = O.determine_interpol_coefficients(f, t, m)
    np.testing.assert_array_equal(a1, o1)
    np.testing.assert_array_equal(a2, o2)
    np.testing.assert_allclose(a1 + a2, 1.0)


def test_gather_triplets_stacks_from_to_between():
    """sampling.gather_triplets = np.vstack of the three slices (data4d_simple.py:210-212) for a batch of triplets."""
    from superresolution_aniso_mri_b200 import sampling
    vol = torch.arange(12, dtype=torch.float32)[:, None, None].expand(12, 5, 7).contiguous()
    rs1, rs2 = np.random.RandomState(17), np.random.RandomState(17)
    trip = [sampling.sample_triplet(z, 12, rs1, kind="brain", slice_selection="adjacent_plus", downsample_steps=4)
            for z in (0, 3, 5, 11)]
    want = [O.sample_triplet(z, 12, rs2, kind="brain", slice_selection="adjacent_plus", downsample_steps=4)
            for z in (0, 3, 5, 11)]
    batch = sampling.gather_triplets(vol, trip)
    assert batch["image"].shape == (4, 3, 5, 7)
    for b, t in enumerate(want):
        assert batch["image"][b, :, 0, 0].tolist() == [t["slice_idx_from"], t["slice_idx_to"], t["inbetween_slice_id"]]
        assert float(batch["alpha_from"][b, 0]) == float(t["alpha_from"]) and float(batch["alpha_to"][b, 0]) == float(t["alpha_to"])
        assert abs(float(batch["alpha_from"][b, 0]) + float(batch["alpha_to"][b, 0]) - 1.0) < 1e-6
    assert batch["is_inbetween"].tolist() == [1.0] * 4


def test_get_random_adjacent_slice_edges():
    """datasets/common.py:34-43: no draw at the volume ends, one rs.choice in the interior."""
    from superresolution_aniso_mri_b200 import sampling
    rs = np.random.RandomState(3)
    state = rs.get_state()[1].copy()
    assert sampling.get_random_adjacent_slice(0, 10, rs, step=2) == 2
    assert sampling.get_random_adjacent_slice(9, 10, rs, step=2) == 7
    assert sampling.get_random_adjacent_slice(1, 10, rs, step=2) == 3
    assert np.array_equal(rs.get_state()[1], state)                     # no random draw so far
    rs_o = np.random.RandomState(3)
    assert sampling.get_random_adjacent_slice(5, 10, rs, step=2) == O.get_random_adjacent_slice(5, 10, rs_o, step=2)
    assert not np.array_equal(rs.get_state()[1], state) or rs.get_state()[2] != 624
