"""CPU tier: host-side logic of the data path that needs no kernel -- batch layout, interpolation coefficients, triplet
gathering -- against the oracle restatements (themselves pinned against the reference in test_oracle_golden.py)."""
import os

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


def test_lpips_weights_are_never_silently_random(tmp_path, monkeypatch):
    """PerceptualLoss must not fall back to an unseeded random VGG16.  Resolution order: explicit tensors, a
    checkpoint path, torchvision's hub file, an explicit seeded opt-in -- else a loud error."""
    import torch
    from superresolution_aniso_mri_b200 import lpips_b200 as L
    monkeypatch.delenv("AESR_VGG16_WEIGHTS", raising=False)
    monkeypatch.delenv("AESR_LPIPS_RANDOM_INIT_SEED", raising=False)
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path / "hub"))
    with pytest.raises(RuntimeError, match="no VGG16 weights"):
        L.resolve_vgg16_state()
    a = L.resolve_vgg16_state(random_init_seed=3)
    b = L.resolve_vgg16_state(random_init_seed=3)
    assert len(a) == 26 and all(torch.equal(x, y) for x, y in zip(a, b))          # identical on every rank
    # a torchvision-style checkpoint (features.N.weight / bias keys) round-trips through vgg_weights=
    idx = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
    sd = {}
    for k, i in enumerate(idx):
        sd["features.%d.weight" % i], sd["features.%d.bias" % i] = a[2 * k], a[2 * k + 1]
    sd["classifier.0.weight"] = torch.zeros(4, 4)
    path = str(tmp_path / "vgg16.pth")
    torch.save(sd, path)
    c = L.resolve_vgg16_state(vgg_weights=path)
    assert all(torch.equal(x, y) for x, y in zip(a, c))
    os.makedirs(str(tmp_path / "hub" / "checkpoints"))
    torch.save(sd, str(tmp_path / "hub" / "checkpoints" / L.VGG16_HUB_FILES[0]))
    d = L.resolve_vgg16_state()                                                   # found in the hub cache
    assert all(torch.equal(x, y) for x, y in zip(a, d))
    rng = torch.random.get_rng_state()
    L.resolve_vgg16_state(random_init_seed=5)
    assert torch.equal(rng, torch.random.get_rng_state())                         # the caller's RNG stream is untouched


def test_bench_flop_model_matches_baseline_tables():
    """bench.py's per-layer MAC model reproduces BASELINE.md / SURVEY.md 8(a): 0.772 / 0.382 GMAC per image at 128^2,
    2.26 / 1.13 at 220^2 (OASIS evaluation), 3.06 / 1.53 at 256^2 (dHCP)."""
    import bench
    for size, enc_g, dec_g in ((128, 0.772, 0.382), (220, 2.26, 1.13), (256, 3.06, 1.53)):
        e, d, _ = bench.layer_macs(size)
        assert abs(sum(e.values()) / 1e9 - enc_g) < 0.006 * enc_g + 0.001, (size, sum(e.values()) / 1e9)
        assert abs(sum(d.values()) / 1e9 - dec_g) < 0.006 * dec_g + 0.001, (size, sum(d.values()) / 1e9)
    enc_b, dec_b = bench.conv_launch_bytes(128)
    assert abs(enc_b / 1024 - 4176.63) < 1.0 and abs(dec_b / 1024 - 1664.0) < 1.0


def test_numa_binding_is_a_noop_without_a_gpu():
    from superresolution_aniso_mri_b200 import parallel
    import torch
    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert parallel.bind_to_gpu_numa(0) is None and os.sched_getaffinity(0) == before
