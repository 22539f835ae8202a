"""GPU tier: evaluation / data-path kernels (SSIM, PSNR, percentile normalisation, pad/crop) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import aesr_oracle as O

pytestmark = pytest.mark.gpu


def test_ssim_psnr_per_slice_matches_oracle(cuda_lib):
    from superresolution_aniso_mri_b200 import evaluation as E
    rng = np.random.RandomState(0)
    vol = O.smooth_phantom(6, 128, seed=3)[:, 0].numpy()
    test = np.clip(vol + 0.03 * rng.randn(*vol.shape).astype(np.float32), 0, 1).astype(np.float32)
    test[2] -= 0.5                                                     # a slice with negative values in `true` role below
    for data_range in (2.0, 1.0):
        ssim, psnr = E.ssim_psnr_slices(vol, test, data_range=data_range)
        for z in range(6):
            assert abs(ssim[z] - O.ssim_slice(vol[z], test[z], data_range=data_range)) < 1e-6       # spec: 1e-3
            assert abs(psnr[z] - O.psnr_slice(vol[z], test[z])) < 1e-4                               # spec: 0.05 dB
    _, psnr_neg = E.ssim_psnr_slices(test, vol)                        # min(true) < 0 on slice 2 -> data_range 2
    assert abs(psnr_neg[2] - O.psnr_slice(test[2], vol[2])) < 1e-4
    assert abs(psnr_neg[0] - O.psnr_slice(test[0], vol[0])) < 1e-4
    # identical images: SSIM 1, PSNR inf (dropped by the batch wrapper)
    s, p = E.ssim_psnr_slices(vol, vol)
    assert np.allclose(s, 1.0, atol=1e-12) and np.all(np.isinf(p))
    # odd sizes
    a = rng.rand(3, 37, 53).astype(np.float32)
    b = rng.rand(3, 37, 53).astype(np.float32)
    s, p = E.ssim_psnr_slices(a, b)
    assert abs(s[1] - O.ssim_slice(a[1], b[1])) < 1e-6 and abs(p[1] - O.psnr_slice(a[1], b[1])) < 1e-4


def test_batch_wrappers_skip_original_slices(cuda_lib):
    from evaluate.metrics import compute_psnr_for_batch, compute_ssim_for_batch
    rng = np.random.RandomState(1)
    ref = rng.rand(11, 64, 64).astype(np.float32)
    rec = np.clip(ref + 0.05 * rng.randn(11, 64, 64).astype(np.float32), 0, 1)
    for d in (None, 2, 3):
        assert abs(compute_ssim_for_batch(ref, rec, downsample_steps=d) -
                   O.compute_ssim_for_batch(ref, rec, downsample_steps=d)) < 1e-6
        assert abs(compute_psnr_for_batch(ref, rec, downsample_steps=d) -
                   O.compute_psnr_for_batch(ref, rec, downsample_steps=d)) < 1e-4


@pytest.mark.parametrize("shape,perc", [((7, 40, 40), (1, 99)), ((10, 128, 128), (1, 99)), ((3, 33, 17), (0, 100)),
                                        ((1, 5, 5), (1, 99))])
def test_percentile_normalize_is_bit_exact_with_numpy(cuda_lib, shape, perc, golden):
    from superresolution_aniso_mri_b200 import evaluation as E
    rng = np.random.RandomState(5 if shape == (7, 40, 40) else 9)
    if shape == (7, 40, 40):
        rng.rand(3, 150, 141); [rng.rand(3, 160, 160) for _ in range(5)]       # same stream position as the golden
    vol = (rng.rand(*shape) * 900 - 50).astype(np.float32)
    if shape != (7, 40, 40):
        vol.flat[::7] = vol.flat[3]                                    # ties around order statistics
    out, lo_hi = E.normalize_img(vol, perc, return_percentiles=True)
    want_lo, want_hi = np.percentile(vol, perc)
    got = lo_hi.cpu().numpy()
    assert got[0] == want_lo and got[1] == want_hi                     # float64, bit-exact
    want = O.normalize_img(vol, perc) if perc == (1, 99) else O.rescale_intensities(vol, perc)
    np.testing.assert_array_equal(out.cpu().numpy(), want.astype(np.float32))
    if shape == (7, 40, 40):
        g = golden("host_logic.npz")                                   # reference normalize_img output
        np.testing.assert_array_equal(out.cpu().numpy().reshape(shape)[:, ::4, ::4], g["normalize_out_sub"].astype(np.float32))


def test_pad_crop_transforms(cuda_lib):
    from superresolution_aniso_mri_b200 import evaluation as E
    rng = np.random.RandomState(5)
    img = rng.rand(4, 3, 150, 141).astype(np.float32)
    pad = E.adjust_to_patch_size(img, 160).cpu().numpy()
    for b in range(4):
        np.testing.assert_array_equal(pad[b], O.adjust_to_patch_size(img[b], 160))
    big = np.pad(img, ((0, 0), (0, 0), (5, 5), (10, 9)))
    cc = E.center_crop(big, 128).cpu().numpy()
    np.testing.assert_array_equal(cc[1], O.center_crop(big[1], 128))
    rs1, rs2 = np.random.RandomState(77), np.random.RandomState(77)
    sq = rng.rand(4, 3, 160, 160).astype(np.float32)
    rc = E.random_crop(sq, 128, rs1).cpu().numpy()
    for b in range(4):
        top, left = O.random_crop_offsets(rs2, 160, 160, 128)
        np.testing.assert_array_equal(rc[b], sq[b, :, top:top + 128, left:left + 128])
