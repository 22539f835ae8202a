"""CPU tier: the oracle (oracle/aesr_oracle.py) against the golden vectors that oracle/make_golden.py produced by
running the UNMODIFIED reference.  Bit-exact unless stated."""
import numpy as np
import pytest
import torch

from oracle import aesr_oracle as O

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def test_init_state_matches_reference_initializer(golden):
    g = golden("init_pins.npz")
    for lw in (32, 16):
        st = O.init_state(O.default_args(128, lw), seed=892372)
        keys = [k for k in st if st[k].dtype.is_floating_point]
        assert keys == list(g["keys_lw%d" % lw])
        assert [str(tuple(st[k].shape)) for k in keys] == list(g["shapes_lw%d" % lw])
        np.testing.assert_array_equal(np.array([st[k].double().sum().item() for k in keys]), g["sum_lw%d" % lw])
        np.testing.assert_array_equal(np.array([st[k].double().abs().sum().item() for k in keys]), g["abs_lw%d" % lw])
    n_params = sum(v.numel() for k, v in O.init_state(O.default_args(128, 32), 892372).items()
                   if k.endswith(("weight", "bias")))
    assert n_params == 443777


def test_small_volume_full_tensors(golden):
    g = golden("infer_small.npz")
    args = O.default_args(64, 16)
    st = O.calibrated_state(args)
    vol = 0.8 * O.smooth_phantom(4, 64, seed=2) + 0.2 * O.synthetic_volume(4, 64, seed=1)
    ar = O.alpha_range_for(2)
    np.testing.assert_array_equal(ar, g["alpha_range"])
    with torch.no_grad():
        z = O.encode(st, args, vol)
        rec = O.decode(st, args, z)
    np.testing.assert_array_equal(z.numpy(), g["z"])
    np.testing.assert_array_equal(rec.numpy(), g["recon"])
    np.testing.assert_array_equal(O.create_super_volume(st, args, vol, ar, use_original=True).numpy(), g["hr"])
    np.testing.assert_array_equal(O.create_super_volume(st, args, vol, ar, use_original=False).numpy(), g["hr_recon"])
    assert float(z.std()) > 0.1          # the calibrated checkpoint is not a vacuous (all ~0) network


def test_eval_twin_with_slice_dropping(golden):
    g = golden("infer_eval_twin.npz")
    args = O.default_args(64, 16)
    st = O.calibrated_state(args)
    vol11 = (0.8 * O.smooth_phantom(11, 64, seed=4) + 0.2 * O.synthetic_volume(11, 64, seed=3))[:, 0]
    out = O.create_super_volume_eval(st, args, vol11, O.alpha_range_for(2), use_original=False, downsample_steps=3,
                                     generate_inbetween_slices=True)
    assert out.shape[0] == 11            # 10 slices of pairs rebuilt + 1 trimmed tail slice re-appended
    np.testing.assert_array_equal(out.numpy(), g["hr"])


@pytest.mark.parametrize("tag,vname,ni", [("cal", "phantom", 6), ("cal", "uniform", 1), ("rnd", "uniform", 6)])
def test_acdc_config1_volume(golden, tag, vname, ni):
    g = golden("infer_acdc.npz")
    args = O.default_args(128, 32)
    st = O.calibrated_state(args) if tag == "cal" else O.init_state(args, seed=892372)
    vol = O.synthetic_volume(10, 128, seed=1) if vname == "uniform" else O.smooth_phantom(10, 128, seed=2)
    hr = O.create_super_volume(st, args, vol, O.alpha_range_for(ni), use_original=True)
    key = "%s_%s_ni%d" % (tag, vname, ni)
    assert hr.shape == (9 * (ni + 1) + 1, 128, 128)
    np.testing.assert_array_equal(hr[:, ::4, ::4].numpy(), g[key + "_sub"])
    np.testing.assert_array_equal(hr.double().sum(dim=(1, 2)).numpy(), g[key + "_slice_sum"])


def test_scales3_readme_literal(golden):
    g = golden("infer_acdc.npz")
    args = O.default_args(128, 16)
    st = O.calibrated_state(args)
    hr = O.create_super_volume(st, args, O.smooth_phantom(5, 128, seed=2), O.alpha_range_for(3), True)
    np.testing.assert_array_equal(hr[:, ::4, ::4].numpy(), g["cal_lw16_phantom_ni3_sub"])


@pytest.mark.parametrize("trainer", ["cardiac", "brain", "plain"])
def test_train_step_small(golden, trainer):
    g = golden("train_small.npz")
    oargs = O.default_args(32, 8)
    st = O.init_state(oargs, seed=892372)
    vgg = O.init_vgg(3)
    lins = load_lins()
    adam = O.AdamState(st, lr=1e-5)
    gen = torch.Generator().manual_seed(11)
    logs = {"loss_ae": [], "loss_ae_dist": [], "loss_ae_dist_extra": [], "loss_latent_1": []}
    for step in range(4):
        img = torch.rand(8, 1, 32, 32, generator=gen)
        sb = torch.rand(4, 1, 32, 32, generator=gen)
        af = at = None
        if trainer == "brain":
            af = torch.tensor([[0.25], [0.5], [0.75], [0.5]])
            at = 1 - af
        lg = O.train_step(st, oargs, adam, img, sb, vgg, lins, ex_loss_weight=0.05, alpha_from=af, alpha_to=at,
                          combined=(trainer != "plain"))
        for k in logs:
            if k in lg:
                logs[k].append(lg[k])
    for k, v in logs.items():
        if v:
            np.testing.assert_array_equal(np.array(v), g["%s_%s" % (trainer, k)])
    keys = [k for k in st if st[k].dtype.is_floating_point]
    assert keys == list(g["%s_state_keys" % trainer])
    np.testing.assert_array_equal(np.array([st[k].double().sum().item() for k in keys]), g["%s_state_sum" % trainer])
    assert int(st["enc.5.num_batches_tracked"]) == (8 if trainer != "plain" else 4)   # 2 BN passes / step (App. B 7)


def load_lins():
    import os
    from superresolution_aniso_mri_b200 import __file__ as pkg
    d = np.load(os.path.join(os.path.dirname(pkg), "data", "lpips_vgg_lin_v0_1.npz"))
    return [torch.from_numpy(d["lin%d" % i]) for i in range(5)]


def test_lpips_forward(golden):
    g = golden("lpips_pins.npz")
    vgg = O.init_vgg(3)
    assert vgg[0][0].double().sum().item() == float(g["vgg_w0_sum"])
    assert vgg[12][0].double().sum().item() == float(g["vgg_w12_sum"])
    gen = torch.Generator().manual_seed(21)
    a = torch.rand(3, 1, 64, 64, generator=gen)
    b = (a + 0.1 * torch.randn(3, 1, 64, 64, generator=gen)).clamp(0, 1)
    with torch.no_grad():
        val = O.lpips_forward(vgg, load_lins(), a, b, normalize=True)
    np.testing.assert_array_equal(val.numpy(), g["lpips"])
    lins = load_lins()
    assert sum(l.numel() for l in lins) == 1472 and min(float(l.min()) for l in lins) >= 0


def test_host_logic(golden):
    g = golden("host_logic.npz")
    rng = np.random.RandomState(5)
    img = rng.rand(3, 150, 141).astype(np.float32)
    np.testing.assert_array_equal(O.adjust_to_patch_size(img, 160)[:, ::5, ::5], g["adjust_160"])
    if "center_128" in g.files:
        np.testing.assert_array_equal(O.center_crop(np.pad(img, ((0, 0), (5, 5), (10, 9))), 128)[:, ::4, ::4],
                                      g["center_128"])
    rs = np.random.RandomState(77)
    offs = []
    for _ in range(5):
        rng.rand(3, 160, 160)
        offs.append(O.random_crop_offsets(rs, 160, 160, 128))
    np.testing.assert_array_equal(np.array(offs), g["random_crop_offsets_seed77"])
    vol = (rng.rand(7, 40, 40) * 900 - 50).astype(np.float32)
    n = O.normalize_img(vol)
    assert str(n.dtype) == str(g["normalize_dtype"])
    np.testing.assert_array_equal(n[:, ::4, ::4], g["normalize_out_sub"])
    af, at = O.determine_interpol_coefficients(np.array([3, 10, 8]), np.array([7, 6, 12]), np.array([4, 8, 11]))
    np.testing.assert_array_equal(af, g["alpha_from"])
    np.testing.assert_array_equal(at, g["alpha_to"])
    for n_sl, d in ((10, 2), (11, 3), (34, 6), (202, 6), (9, 4)):
        np.testing.assert_array_equal(O.synth_slice_mask(n_sl, d)[1], g["smask_%d_%d" % (n_sl, d)])
        np.testing.assert_array_equal(O.determine_original_sliceids(n_sl, d), g["origids_%d_%d" % (n_sl, d)])
    for ni in (1, 2, 3, 5, 6):
        hi, lo = O.interp_weights(O.alpha_range_for(ni))
        np.testing.assert_array_equal(hi, g["w_hi_ni%d" % ni])
        np.testing.assert_array_equal(lo, g["w_lo_ni%d" % ni])
    # the tempting shortcut 1 - float32(alpha) is NOT what torch computes (SURVEY.md section 7 hard part 5)
    hi, lo = O.interp_weights(O.alpha_range_for(6))
    assert np.any((np.float32(1) - hi) != lo)


AUGMENT_CASES = (("acdc", dict(width=128, aug_patch=160, center=True, intensity_first=True)),
                 ("acdc_small", dict(width=64, aug_patch=96, center=True, intensity_first=True)),
                 ("oasis", dict(width=64, aug_patch=220, center=False, intensity_first=False)),
                 ("dhcp_crop", dict(width=128, aug_patch=None, center=False, intensity_first=False)))


def test_augment_chain_against_reference_golden(golden):
    """The training transform chains (AdjustToPatchSize / CenterCrop / RandomCrop / RandomIntensity / RandomRotation,
    datasets/shared_transforms.py) in the ACDC and the brain order: values and RandomState draw order pinned against
    the reference's own classes (oracle/make_golden.py::gold_augment)."""
    g = golden("augment_pins.npz")
    rng = np.random.RandomState(int(g["input_seed"]))
    for name, kw in AUGMENT_CASES:
        rs = np.random.RandomState(int(g["rs_seed"]))
        shape = tuple(int(v) for v in g["%s_shape" % name])
        for i in range(4):
            img = rng.rand(*shape).astype(np.float32)
            got, d = O.augment_sample(img, rs, **kw)
            assert got.shape == (shape[0], kw["width"], kw["width"]) and got.dtype == np.float32
            np.testing.assert_array_equal(got[:, ::7, ::5], g["%s_out%d" % (name, i)])
            np.testing.assert_array_equal(np.array([d["top"], d["left"], d["gain"], d["cutoff"], d["k"]], dtype=np.float64),
                                          g["%s_draws" % name][i])


def test_thick_slices_against_reference_golden(golden):
    """datasets/common_brains.py::simulate_thick_slices (scipy gaussian_filter1d per column), incl. a filter radius
    larger than the volume."""
    g = golden("thick_slices_pins.npz")
    rs = np.random.RandomState(int(g["seed"]))
    for k in range(5):
        cfg = g["cfg%d" % k]
        vol = rs.rand(*[int(v) for v in cfg[:3]]).astype(np.float32)
        np.testing.assert_array_equal(O.simulate_thick_slices(vol, float(cfg[3])), g["thick%d" % k])


def test_triplet_sampling_against_reference_golden(golden):
    """Dataset __getitem__ index sampling (datasets/common.py:34-43, data4d_simple.py:191-205, common_brains.py:241-260):
    oracle and the product's host function against the sequences the reference's own methods produced."""
    from superresolution_aniso_mri_b200 import sampling
    g = golden("sampling_pins.npz")
    for key in g.files:
        if key == "seed":
            continue
        kind, sel = key.split("_")[0], "_".join(key.split("_")[1:-2])
        ds, Z = int(key.split("_")[-2]), int(key.split("_")[-1])
        for fn in (O.sample_triplet, sampling.sample_triplet):
            rs = np.random.RandomState(int(g["seed"]))
            rows = g[key]
            r = 0
            for rep in range(3):
                for z in range(Z):
                    if r >= len(rows) or int(rows[r][0]) != z:
                        # the reference raised (empty open interval) or the generator skipped this slice: replay the draws
                        if kind == "brain" and sel == "mix":
                            try:
                                fn(z, Z, rs, kind=kind, slice_selection=sel, downsample_steps=ds)
                            except ValueError:
                                pass
                        continue
                    t = fn(z, Z, rs, kind=kind, slice_selection=sel, downsample_steps=ds)
                    got = [z, t["slice_idx_from"], t["slice_idx_to"], t["inbetween_slice_id"], float(t["is_inbetween"]),
                           float(t["alpha_from"]), float(t["alpha_to"])]
                    assert got == list(rows[r]), (key, fn.__module__, rep, z)
                    r += 1
            assert r == len(rows)


def test_ssim_psnr_properties():
    """scikit-image is absent and unpinned (parity unpinned): check the restatement on analytic properties."""
    rng = np.random.RandomState(0)
    a = rng.rand(64, 64).astype(np.float32)
    assert abs(O.ssim_slice(a, a) - 1.0) < 1e-12
    b = np.clip(a + 0.05 * rng.randn(64, 64).astype(np.float32), 0, 1)
    s = O.ssim_slice(a, b)
    assert 0 < s < 1 and abs(O.ssim_slice(b, a) - s) < 1e-12
    assert abs(O.psnr_slice(a, b) - 10 * np.log10(1.0 / np.mean((a.astype(np.float64) - b) ** 2))) < 1e-9
    # uniform filter restatement against a direct window mean
    f = O._uniform_filter_reflect(a.astype(np.float64), 7)
    p = np.pad(a.astype(np.float64), 3, mode="symmetric")
    assert abs(f[10, 20] - p[10:17, 20:27].mean()) < 1e-12 and abs(f[0, 0] - p[0:7, 0:7].mean()) < 1e-12
    # ... and against scipy.ndimage.uniform_filter itself: the one numerical component scikit-image's structural_similarity
    # delegates to (its own code is the ~15 lines of arithmetic ssim_slice restates), on even / odd / tiny sizes
    import scipy.ndimage
    for shape in ((64, 64), (37, 53), (9, 8), (7, 7), (220, 220)):
        img = rng.rand(*shape)
        np.testing.assert_allclose(O._uniform_filter_reflect(img, 7), scipy.ndimage.uniform_filter(img, size=7), rtol=0, atol=1e-13)
    # the whole metric with scipy's filter in place of the restated one
    x, y = a.astype(np.float64), b.astype(np.float64)
    uf = lambda v: scipy.ndimage.uniform_filter(v, size=7)      # noqa: E731
    ux, uy, uxx, uyy, uxy = uf(x), uf(y), uf(x * x), uf(y * y), uf(x * y)
    cn = 49.0 / 48.0
    vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
    C1, C2 = (0.01 * 2.0) ** 2, (0.03 * 2.0) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    assert abs(S[3:-3, 3:-3].mean() - s) < 1e-12


def test_vif_oracle_against_reference_golden(golden):
    """tests/golden/vif_pins.npz holds outputs of the REFERENCE's evaluate/vifvec.py::vifp_mscale and
    evaluate/metrics.py::compute_vif_for_batch (oracle/make_golden.py::gold_vif, which also asserts every uint8 plane of
    the restated gaussian filter equal to scipy.ndimage's): the oracle reproduces them."""
    import scipy.ndimage
    g = golden("vif_pins.npz")
    rs = np.random.RandomState(11)
    for (h, w) in ((37, 53), (128, 128), (16, 9)):
        a = rs.randint(0, 256, size=(h, w)).astype(np.uint8)
        a[: h // 3, : w // 2] = 100
        for sd in (3.4, 1.8, 1.0, 0.6):
            np.testing.assert_array_equal(O.gaussian_filter_u8(a, sd), scipy.ndimage.gaussian_filter(a, sd))
    vol = O.smooth_phantom(6, 128, seed=2)[:, 0].numpy()
    noise = rs.normal(0, 0.05, vol.shape).astype(np.float32)
    noisy = np.clip(vol + noise, 0, 1).astype(np.float32)
    for z in (0, 3, 5):
        v = O.vifp_mscale_u8(O.quantize_u8(vol[z]), O.quantize_u8(noisy[z]))
        assert abs(v - g["vif_noisy"][z]) < 1e-12
    with np.errstate(divide="ignore", invalid="ignore"):
        assert abs(O.vifp_mscale_u8(O.quantize_u8(vol[1]), O.quantize_u8(vol[1])) - g["vif_same"][1]) < 1e-12
        assert np.isnan(O.vifp_mscale_u8(O.quantize_u8(vol[1]), O.quantize_u8(np.zeros_like(vol[1])))) == np.isnan(g["vif_black"][1])
    for ds in (None, 2, 3):
        assert abs(O.compute_vif_for_batch(vol, noisy, downsample_steps=ds) - float(g["vif_batch_ds%s" % ds])) < 1e-12
    m = O.compute_metrics(vol, noisy, 2)
    assert set(m) == {"ssim", "psnr", "vif", "ssim_synth", "psnr_synth", "vif_synth", "ssim_recon", "psnr_recon", "vif_recon"}
    assert O.determine_last_slice(11, 3) == 9 and O.determine_last_slice(10, 3) == 9


def test_reference_trained_checkpoint_pins(golden):
    """tests/golden/trained_ckpt.npz: a checkpoint trained by the reference itself (make_golden.gold_trained) and the
    reference's synthesis outputs for it; the oracle reproduces them bit for bit, and the checkpoint is a real model
    (it reconstructs held-out phantoms, BN statistics are trained, latents are O(1))."""
    from collections import OrderedDict
    g = golden("trained_ckpt.npz")
    st = OrderedDict((k[len("state__"):], torch.from_numpy(g[k].copy())) for k in g.files if k.startswith("state__"))
    args = O.default_args(64, 16)
    assert list(st.keys()) == list(O.init_state(args, seed=1).keys())
    assert int(st["enc.5.num_batches_tracked"]) == 2 * int(g["steps"])          # enc(x) and enc(slice_between) per step
    vol = O.mri_phantom(10, 128, seed=41)
    hr = O.create_super_volume(st, args, vol, O.alpha_range_for(2), use_original=True)
    np.testing.assert_array_equal(hr[:, ::4, ::4].numpy(), g["acdc128_ni2_sub"])
    np.testing.assert_array_equal(hr.double().sum(dim=(1, 2)).numpy(), g["acdc128_ni2_slice_sum"])
    v3 = O.mri_phantom(9, 220, seed=47)[:, 0]
    out = O.create_super_volume_eval(st, args, v3, O.alpha_range_for(3), use_original=False, downsample_steps=4,
                                     generate_inbetween_slices=True)
    np.testing.assert_array_equal(out[:, ::5, ::5].numpy(), g["oasis220_ds4_sub"])
    with torch.no_grad():
        z = O.encode(st, args, vol)
        rec = O.decode(st, args, z)
    assert float(torch.mean((rec - vol) ** 2)) < 3e-3 and float(z.std()) > 0.05
    assert abs(float(torch.mean((rec - vol) ** 2)) - float(g["acdc128_recon_mse"])) < 1e-9
