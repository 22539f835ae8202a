"""Pin the oracle against the UNMODIFIED reference and write the golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden [--only NAME ...]

Every fixture is produced by the reference's own code (imported through oracle/ref_bootstrap.py) on seeded
synthetic inputs; the same run asserts that oracle/aesr_oracle.py reproduces it (bit-exact unless noted).
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

from oracle import aesr_oracle as O
from oracle import ref_bootstrap as rb

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PKG_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "superresolution_aniso_mri_b200", "data")


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def _ref_model(args, state=None, seed=892372):
    from networks.acai_vanilla import VanillaACAI
    torch.manual_seed(seed)
    m = VanillaACAI(dict(args))
    if state is not None:
        m.load_state_dict(state)
    return m.eval()


def gold_init():
    out = {}
    for lw in (32, 16):
        args = O.default_args(128, lw)
        sd = _ref_model(args).state_dict()
        st = O.init_state(args, seed=892372)
        assert list(sd.keys()) == list(st.keys())
        assert all(torch.equal(sd[k], st[k]) for k in sd), "init_state != reference Initializer"
        keys = [k for k in sd if sd[k].dtype.is_floating_point]
        out["keys_lw%d" % lw] = np.array(keys)
        out["sum_lw%d" % lw] = np.array([sd[k].double().sum().item() for k in keys])
        out["abs_lw%d" % lw] = np.array([sd[k].double().abs().sum().item() for k in keys])
        out["shapes_lw%d" % lw] = np.array([str(tuple(sd[k].shape)) for k in keys])
    _save("init_pins.npz", **out)


def gold_lpips_lin():
    tr_args = rb.reference_args(width=32, latent_width=8, batch_size=2)
    tr = rb.make_reference_trainer(dict(tr_args))
    net = tr.percept_criterion.model.net
    lins = {"lin%d" % i: getattr(net, "lin%d" % i).model[1].weight.data.numpy().copy() for i in range(5)}
    os.makedirs(PKG_DATA, exist_ok=True)
    np.savez(os.path.join(PKG_DATA, "lpips_vgg_lin_v0_1.npz"), **lins)
    print("wrote lin heads", {k: v.shape for k, v in lins.items()})
    # LPIPS forward pins: reference PerceptualLoss on seeded inputs
    vgg = O.init_vgg(3)
    ws = [v for k, v in net.net.state_dict().items() if k.endswith("weight")]
    assert all(torch.equal(a, b[0]) for a, b in zip(ws, vgg)), "init_vgg != torchvision vgg16 init"
    g = torch.Generator().manual_seed(21)
    a = torch.rand(3, 1, 64, 64, generator=g)
    b = (a + 0.1 * torch.randn(3, 1, 64, 64, generator=g)).clamp(0, 1)
    with torch.no_grad():
        ref = tr.percept_criterion(a, b, normalize=True)
        mine = O.lpips_forward(vgg, [torch.from_numpy(lins["lin%d" % i]) for i in range(5)], b, a, normalize=True)
        # reference call convention: percept_criterion(reference, synthesized) => forward(pred=reference, target=synth)
        mine2 = O.lpips_forward(vgg, [torch.from_numpy(lins["lin%d" % i]) for i in range(5)], a, b, normalize=True)
    assert torch.equal(ref, mine2) or torch.allclose(ref, mine2, rtol=0, atol=0), (ref.flatten(), mine2.flatten())
    _save("lpips_pins.npz", lpips=ref.numpy(), lpips_swapped=mine.numpy(), vgg_seed=np.array(3), input_seed=np.array(21),
          vgg_w0_sum=np.array(vgg[0][0].double().sum().item()), vgg_w12_sum=np.array(vgg[12][0].double().sum().item()))


def gold_infer():
    ghv = rb.reference_generate_module()
    ec = rb.reference_eval_common_module()
    # --- small config, full tensors -------------------------------------------------------------
    args = O.default_args(64, 16)
    st = O.calibrated_state(args)
    m = _ref_model(args, st)
    vol = 0.8 * O.smooth_phantom(4, 64, seed=2) + 0.2 * O.synthetic_volume(4, 64, seed=1)
    ar = O.alpha_range_for(2)
    with torch.no_grad():
        z = m.encode(vol)
        rec = m.decode(z)
    hr = ghv.create_super_volume(rb.CpuEvalTrainer(m), vol, ar, use_original=True)["upsampled_image"]
    with rb.cuda_to_cpu():
        hr_rec = ghv.create_super_volume(rb.CpuEvalTrainer(m), vol, ar, use_original=False)["upsampled_image"]
    assert torch.equal(hr_rec, O.create_super_volume(st, args, vol, ar, use_original=False))
    assert torch.equal(z, O.encode(st, args, vol)) and torch.equal(rec, O.decode(st, args, z))
    assert torch.equal(hr, O.create_super_volume(st, args, vol, ar, use_original=True))
    _save("infer_small.npz", z=z.numpy(), recon=rec.numpy(), hr=hr.numpy(), hr_recon=hr_rec.numpy(), alpha_range=ar)
    # --- evaluation twin with slice dropping ------------------------------------------------------
    vol11 = (0.8 * O.smooth_phantom(11, 64, seed=4) + 0.2 * O.synthetic_volume(11, 64, seed=3))[:, 0]
    with rb.cuda_to_cpu():
        out = ec.create_super_volume(rb.CpuEvalTrainer(m), vol11, alpha_range=O.alpha_range_for(2), use_original=False,
                                     downsample_steps=3, generate_inbetween_slices=True)["upsampled_image"]
    mine = O.create_super_volume_eval(st, args, vol11, O.alpha_range_for(2), use_original=False, downsample_steps=3,
                                      generate_inbetween_slices=True)
    assert out.shape == mine.shape and torch.equal(out, mine), (out.shape, mine.shape)
    _save("infer_eval_twin.npz", hr=out.numpy())
    # --- ACDC config 1: Z=10, 128^2, ni=6 and ni=1, calibrated and literal random init -------------
    args = O.default_args(128, 32)
    pins = {}
    for tag, state in (("cal", O.calibrated_state(args)), ("rnd", O.init_state(args, seed=892372))):
        m = _ref_model(args, state)
        for vname, vol in (("uniform", O.synthetic_volume(10, 128, seed=1)),
                           ("phantom", O.smooth_phantom(10, 128, seed=2))):
            for ni in (6, 1):
                ar = O.alpha_range_for(ni)
                t0 = time.time()
                hr = ghv.create_super_volume(rb.CpuEvalTrainer(m), vol, ar, use_original=True)["upsampled_image"]
                dt = time.time() - t0
                mine = O.create_super_volume(state, args, vol, ar, use_original=True)
                assert torch.equal(hr, mine)
                key = "%s_%s_ni%d" % (tag, vname, ni)
                pins[key + "_sub"] = hr[:, ::4, ::4].numpy()
                pins[key + "_slice_sum"] = hr.double().sum(dim=(1, 2)).numpy()
                print(key, tuple(hr.shape), "ref cpu %.2fs -> %.1f slices/s" % (dt, 9 * ni / dt))
        with torch.no_grad():
            z = m.encode(O.smooth_phantom(10, 128, seed=2))
        pins["%s_phantom_z_sub" % tag] = z[:, ::8, ::4, ::4].numpy()
    # scales=3 (README-literal latent_width=16)
    args3 = O.default_args(128, 16)
    st3 = O.calibrated_state(args3)
    m3 = _ref_model(args3, st3)
    hr = ghv.create_super_volume(rb.CpuEvalTrainer(m3), O.smooth_phantom(5, 128, seed=2), O.alpha_range_for(3),
                                 use_original=True)["upsampled_image"]
    assert torch.equal(hr, O.create_super_volume(st3, args3, O.smooth_phantom(5, 128, seed=2), O.alpha_range_for(3), True))
    pins["cal_lw16_phantom_ni3_sub"] = hr[:, ::4, ::4].numpy()
    pins["cal_lw16_phantom_ni3_slice_sum"] = hr.double().sum(dim=(1, 2)).numpy()
    _save("infer_acdc.npz", **pins)


def gold_train_small():
    out = {}
    for trainer in ("cardiac", "brain", "plain"):
        args = rb.reference_args(dataset="ACDC" if trainer != "brain" else "dHCP", width=32, latent_width=8,
                                 batch_size=4, ex_loss_weight1=0.05)
        tr = rb.make_reference_trainer(dict(args), trainer=trainer)
        oargs = O.default_args(32, 8)
        st = O.init_state(oargs, seed=892372)
        vgg = O.init_vgg(3)
        net = tr.percept_criterion.model.net if tr.percept_criterion is not None else None
        lins = [getattr(net, "lin%d" % i).model[1].weight.data.clone() for i in range(5)] if net is not None else None
        adam = O.AdamState(st, lr=args["lr"])
        g = torch.Generator().manual_seed(11)
        af = at = None
        for step in range(4):
            img = torch.rand(8, 1, 32, 32, generator=g)
            sb = torch.rand(4, 1, 32, 32, generator=g)
            b = {"image": img, "slice_between": sb}
            if trainer == "brain":
                af = torch.tensor([[0.25], [0.5], [0.75], [0.5]])
                at = 1 - af
                b["alpha_from"], b["alpha_to"] = af, at
            tr.train(b, keep_predictions=False)
            lg = O.train_step(st, oargs, adam, img, sb, vgg, lins, ex_loss_weight=0.05, alpha_from=af, alpha_to=at,
                              combined=(trainer != "plain"))
            for k in ("loss_ae", "loss_ae_dist", "loss_ae_dist_extra", "loss_latent_1"):
                if k in tr.losses:
                    assert tr.losses[k][-1] == lg[k], (trainer, step, k, tr.losses[k][-1], lg[k])
        sd = tr.model.state_dict()
        assert all(torch.equal(sd[k], st[k]) for k in sd), "post-step state differs (%s)" % trainer
        for k in tr.losses:
            out["%s_%s" % (trainer, k)] = np.array(tr.losses[k])
        keys = [k for k in sd if sd[k].dtype.is_floating_point]
        out["%s_state_sum" % trainer] = np.array([sd[k].double().sum().item() for k in keys])
        out["%s_state_keys" % trainer] = np.array(keys)
        print(trainer, {k: v[-1] for k, v in tr.losses.items()})
    _save("train_small.npz", **out)


def acdc_batch(i: int, B: int = 12, size: int = 128):
    """Config 2 synthetic batch i of the fixed cycle of 8.  Smooth-ish triplets: the
    'between' slice is the mean of from/to plus noise so LPIPS/MSE have structure."""
    g = torch.Generator().manual_seed(1000 + i)
    base = torch.rand(B, 1, size // 8, size // 8, generator=g)
    up = torch.nn.functional.interpolate(base, size=(size, size), mode="bilinear", align_corners=False)
    a = (up + 0.15 * torch.rand(B, 1, size, size, generator=g)).clamp(0, 1)
    b = (up.flip(-1) * 0.5 + up * 0.5 + 0.15 * torch.rand(B, 1, size, size, generator=g)).clamp(0, 1)
    mid = (0.5 * a + 0.5 * b + 0.05 * torch.rand(B, 1, size, size, generator=g)).clamp(0, 1)
    return torch.cat([a, b], dim=0), mid


def gold_train_acdc(steps=200):
    torch.set_num_threads(os.cpu_count())
    args = rb.reference_args(width=128, latent_width=32, batch_size=12, ex_loss_weight1=0.05, lr=1e-5)
    tr = rb.make_reference_trainer(dict(args), trainer="cardiac")
    t0 = time.time()
    for s in range(steps):
        img, mid = acdc_batch(s % 8)
        tr.train({"image": img, "slice_between": mid}, keep_predictions=False)
        if s % 10 == 0:
            print(s, tr.losses["loss_ae"][-1], "%.1fs" % (time.time() - t0), flush=True)
    out = {k: np.array(v) for k, v in tr.losses.items()}
    sd = tr.model.state_dict()
    keys = [k for k in sd if sd[k].dtype.is_floating_point]
    out["state_keys"] = np.array(keys)
    out["state_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    out["sec_per_step"] = np.array((time.time() - t0) / steps)
    out["cores"] = np.array(os.cpu_count())
    _save("train_acdc_%d.npz" % steps, **out)


def trained_batch(i: int, B: int = 8, size: int = 64):
    """Seeded MRI-like triplets for the reference-trained checkpoint: sample b = slices (3b, 3b+1, 3b+2) of a phantom."""
    v = O.mri_phantom(3 * B, size, seed=5000 + i)
    return torch.cat([v[0::3], v[2::3]], dim=0), v[1::3].clone()


def gold_trained(steps=400):
    """A checkpoint TRAINED BY THE REFERENCE (AETrainerEndToEnd.train, lr 1e-3, MSE + 0.05 LPIPS on MRI-like phantoms) so that
    BatchNorm running statistics, gamma / beta and the filters are those of a model that reconstructs images -- and the
    reference's own synthesis outputs for it at BASELINE configs 1 (ACDC 128^2, ni 6), 3 (OASIS 220^2, ds 4) and
    4 / published (dHCP 256^2, ds 4 and 6).  All three share the scales=2 architecture (fully convolutional)."""
    torch.set_num_threads(os.cpu_count())
    args = rb.reference_args(width=64, latent_width=16, batch_size=8, ex_loss_weight1=0.05, lr=1e-3)
    tr = rb.make_reference_trainer(dict(args), trainer="cardiac")
    t0 = time.time()
    for s in range(steps):
        img, mid = trained_batch(s)
        tr.train({"image": img, "slice_between": mid}, keep_predictions=False)
        if s % 20 == 0:
            print(s, tr.losses["loss_ae"][-1], tr.losses["loss_ae_dist"][-1], "%.1fs" % (time.time() - t0), flush=True)
    sd = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    out = {"state__" + k: v.numpy() for k, v in sd.items()}
    out["train_loss_ae"] = np.array(tr.losses["loss_ae"])
    out["train_loss_ae_dist"] = np.array(tr.losses["loss_ae_dist"])
    ghv = rb.reference_generate_module()
    ec = rb.reference_eval_common_module()
    oargs = O.default_args(64, 16)
    m = _ref_model(oargs, sd)
    # config 1: ACDC-shaped volume through generate_hr_volumes.create_super_volume
    vol = O.mri_phantom(10, 128, seed=41)
    for ni in (6, 2):
        ar = O.alpha_range_for(ni)
        hr = ghv.create_super_volume(rb.CpuEvalTrainer(m), vol, ar, use_original=True)["upsampled_image"]
        assert torch.equal(hr, O.create_super_volume(sd, oargs, vol, ar, use_original=True))
        out["acdc128_ni%d_sub" % ni] = hr[:, ::4, ::4].numpy()
        out["acdc128_ni%d_slice_sum" % ni] = hr.double().sum(dim=(1, 2)).numpy()
    with torch.no_grad():
        rec = m(vol)
    out["acdc128_recon_mse"] = np.array(torch.mean((rec - vol) ** 2).item())
    print("reconstruction MSE of the trained model on a held-out phantom: %.5f" % out["acdc128_recon_mse"])
    # configs 3 / 4 / published: the evaluation twin (slice dropping) at 220^2 and 256^2
    for tag, size, Z, ds in (("oasis220_ds4", 220, 9, 4), ("dhcp256_ds4", 256, 9, 4), ("dhcp256_ds6", 256, 13, 6)):
        v3 = O.mri_phantom(Z, size, seed=43 + ds)[:, 0]
        with rb.cuda_to_cpu():
            hr = ec.create_super_volume(rb.CpuEvalTrainer(m), v3, alpha_range=O.alpha_range_for(ds - 1), use_original=False,
                                        downsample_steps=ds, generate_inbetween_slices=True)["upsampled_image"]
        mine = O.create_super_volume_eval(sd, oargs, v3, O.alpha_range_for(ds - 1), use_original=False, downsample_steps=ds,
                                          generate_inbetween_slices=True)
        assert hr.shape == mine.shape and torch.equal(hr, mine), (tag, hr.shape, mine.shape)
        out[tag + "_sub"] = hr[:, ::5, ::5].numpy()
        out[tag + "_slice_sum"] = hr.double().sum(dim=(1, 2)).numpy()
        print(tag, tuple(hr.shape), "synth-vs-truth mean abs %.4f" % (hr - v3[: hr.shape[0]]).abs().mean().item())
    out["steps"], out["sec_per_step"] = np.array(steps), np.array((time.time() - t0) / steps)
    _save("trained_ckpt.npz", **out)


def gold_transforms():
    rb.bootstrap()
    import datasets.shared_transforms as stf
    from datasets.common_brains import determine_interpol_coefficients, prepare_batch_pairs
    from evaluate.quantitative_comparison import generate_synth_slices_mask
    from evaluate.metrics import determine_original_sliceids
    import generate_hr_volumes as ghv
    out = {}
    g = np.random.RandomState(5)
    img = g.rand(3, 150, 141).astype(np.float32)
    rs = np.random.RandomState(1234)
    padded = stf.AdjustToPatchSize((160, 160))({"image": img.copy()})["image"] if _sig(stf.AdjustToPatchSize) else None
    if padded is not None:
        assert np.array_equal(padded, O.adjust_to_patch_size(img, 160))
        out["adjust_160"] = padded[:, ::5, ::5]
    try:
        cc = stf.CenterCrop(128)({"image": np.pad(img, ((0, 0), (5, 5), (10, 9)))})["image"]
        assert np.array_equal(cc, O.center_crop(np.pad(img, ((0, 0), (5, 5), (10, 9))), 128))
        out["center_128"] = cc[:, ::4, ::4]
    except Exception as e:                                    # pragma: no cover - reference API rot
        print("CenterCrop skipped:", repr(e))
    rs1, rs2 = np.random.RandomState(77), np.random.RandomState(77)
    rc = stf.RandomCrop(128, rs=rs1)
    offs = []
    for _ in range(5):
        big = g.rand(3, 160, 160).astype(np.float32)
        got = rc({"image": big})["image"]
        top, left = O.random_crop_offsets(rs2, 160, 160, 128)
        assert np.array_equal(got, big[:, top:top + 128, left:left + 128])
        offs.append((top, left))
    out["random_crop_offsets_seed77"] = np.array(offs)
    vol = (g.rand(7, 40, 40) * 900 - 50).astype(np.float32)
    n = ghv.normalize_img(vol)
    assert np.array_equal(n, O.normalize_img(vol))
    out["normalize_in_seed"] = np.array(5)
    out["normalize_out_sub"] = n[:, ::4, ::4]
    out["normalize_dtype"] = np.array(str(n.dtype))
    af, at = determine_interpol_coefficients(np.array([3, 10, 8]), np.array([7, 6, 12]), np.array([4, 8, 11]))
    a2, t2 = O.determine_interpol_coefficients(np.array([3, 10, 8]), np.array([7, 6, 12]), np.array([4, 8, 11]))
    assert np.array_equal(af, a2) and np.array_equal(at, t2)
    out["alpha_from"], out["alpha_to"] = af, at
    b = torch.rand(4, 3, 8, 8, generator=torch.Generator().manual_seed(1))
    d = prepare_batch_pairs({"image": b.clone()})
    mine = O.prepare_batch_pairs(b)
    assert torch.equal(d["image"], mine["image"]) and torch.equal(d["slice_between"], mine["slice_between"])
    for n_sl, d_steps in ((10, 2), (11, 3), (34, 6), (202, 6), (9, 4)):
        r, s = generate_synth_slices_mask(n_sl, d_steps)
        r2, s2 = O.synth_slice_mask(n_sl, d_steps)
        assert np.array_equal(r, r2) and np.array_equal(s, s2)
        ids = determine_original_sliceids(np.zeros((n_sl, 2, 2)), d_steps)
        assert np.array_equal(ids, O.determine_original_sliceids(n_sl, d_steps))
        out["smask_%d_%d" % (n_sl, d_steps)] = s
        out["origids_%d_%d" % (n_sl, d_steps)] = ids
    for ni in (1, 2, 3, 5, 6):
        ar = np.linspace(0, 1, ni + 2, endpoint=True)[1:-1]
        hi, lo = O.interp_weights(ar)
        z1, z2 = torch.full((1,), 1.0), torch.full((1,), 1.0)
        for k, a in enumerate(ar):       # what torch actually multiplies by
            assert (a * z1).item() == float(hi[k]) and ((1 - a) * z2).item() == float(lo[k])
        out["w_hi_ni%d" % ni], out["w_lo_ni%d" % ni] = hi, lo
    _save("host_logic.npz", **out)


def gold_vif():
    """VIF: the oracle restatement must equal the reference's evaluate/vifvec.py::vifp_mscale (scipy.ndimage underneath)
    on uint8 slices bit for bit in every uint8 plane, and to float64 rounding in the final ratio; also pins
    evaluate/metrics.py::compute_vif_for_batch and evaluate/create_HR_images.py::compute_metrics' slice selection."""
    import scipy.ndimage
    from evaluate.vifvec import vifp_mscale
    from evaluate import metrics as M
    rs = np.random.RandomState(11)
    out = {}
    # (i) every intermediate uint8 plane of the gaussian filter, all four sigmas, odd sizes, flat regions included
    for k, (h, w) in enumerate(((37, 53), (128, 128), (16, 9))):
        a = rs.randint(0, 256, size=(h, w)).astype(np.uint8)
        a[: h // 3, : w // 2] = 100                     # flat patch: sum(w) * 100 sits next to an integer boundary
        for sd in (3.4, 1.8, 1.0, 0.6):
            want = scipy.ndimage.gaussian_filter(a, sd)
            got = O.gaussian_filter_u8(a, sd)
            assert want.dtype == np.uint8 and np.array_equal(want, got), (h, w, sd)
    # (ii) the metric itself on phantom-like slices and their degraded versions
    vol = O.smooth_phantom(6, 128, seed=2)[:, 0].numpy()
    noise = rs.normal(0, 0.05, vol.shape).astype(np.float32)
    blurred = scipy.ndimage.gaussian_filter(vol, (0, 1.5, 1.5)).astype(np.float32)
    cases = {"noisy": np.clip(vol + noise, 0, 1).astype(np.float32), "blurred": blurred, "same": vol.copy(),
             "black": np.zeros_like(vol)}
    for name, dist in cases.items():
        ref_u8, dist_u8 = O.quantize_u8(vol), O.quantize_u8(dist)
        vals = []
        for z in range(vol.shape[0]):
            with np.errstate(divide="ignore", invalid="ignore"):
                r = vifp_mscale(ref_u8[z], dist_u8[z])
                m = O.vifp_mscale_u8(ref_u8[z], dist_u8[z])
            assert (np.isnan(r) and np.isnan(m)) or abs(r - m) <= 1e-12 * max(1.0, abs(r)), (name, z, r, m)
            vals.append(r)
        out["vif_" + name] = np.array(vals, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        for ds in (None, 2, 3):
            r = M.compute_vif_for_batch(vol, cases["noisy"], downsample_steps=ds)
            m = O.compute_vif_for_batch(vol, cases["noisy"], downsample_steps=ds)
            assert abs(r - m) <= 1e-12, (ds, r, m)
            out["vif_batch_ds%s" % ds] = np.array(r)
    # 220 x 220 (OASIS eval size), odd size, rectangular
    big = rs.rand(2, 220, 220).astype(np.float32)
    big2 = np.clip(big + rs.normal(0, 0.1, big.shape), 0, 1).astype(np.float32)
    out["vif_220"] = np.array([vifp_mscale(O.quantize_u8(big)[z], O.quantize_u8(big2)[z]) for z in range(2)])
    rect = rs.rand(2, 45, 77).astype(np.float32)
    rect2 = np.clip(rect * 0.8 + 0.1, 0, 1).astype(np.float32)
    out["vif_rect"] = np.array([vifp_mscale(O.quantize_u8(rect)[z], O.quantize_u8(rect2)[z]) for z in range(2)])
    for z in range(2):
        assert abs(out["vif_220"][z] - O.vifp_mscale_u8(O.quantize_u8(big)[z], O.quantize_u8(big2)[z])) < 1e-12
        assert abs(out["vif_rect"][z] - O.vifp_mscale_u8(O.quantize_u8(rect)[z], O.quantize_u8(rect2)[z])) < 1e-12
    out["seed"] = np.array(11)
    _save("vif_pins.npz", **out)


def gold_augment():
    """Training transform chains (datasets/shared_transforms.py) with a shared RandomState: the ACDC order
    (train_cardiac_aesr.py:90-96) and the brain order (datasets/common_brains.py:77-80).  Pins the oracle's
    augment_sample (values AND the RandomState draw order) against the reference classes."""
    rb.bootstrap()
    import datasets.shared_transforms as stf
    out = {}
    g = np.random.RandomState(21)
    cases = (("acdc", dict(width=128, aug_patch=160, center=True, intensity_first=True), (3, 150, 171)),
             ("acdc_small", dict(width=64, aug_patch=96, center=True, intensity_first=True), (3, 70, 101)),
             ("oasis", dict(width=64, aug_patch=220, center=False, intensity_first=False), (3, 176, 208)),
             ("dhcp_crop", dict(width=128, aug_patch=None, center=False, intensity_first=False), (3, 256, 256)))
    for name, kw, shape in cases:
        rs_ref, rs_mine = np.random.RandomState(4321), np.random.RandomState(4321)
        chain = []
        if kw["aug_patch"] is not None:
            chain.append(stf.AdjustToPatchSize((kw["aug_patch"], kw["aug_patch"])))
            if kw["center"]:
                chain.append(stf.CenterCrop((kw["aug_patch"], kw["aug_patch"])))
        chain.append(stf.RandomCrop(kw["width"], rs=rs_ref))
        if kw["intensity_first"]:
            chain += [stf.RandomIntensity(rs=rs_ref, slice_mask=None), stf.RandomRotation(rs_ref)]
        else:
            chain += [stf.RandomRotation(rs=rs_ref), stf.RandomIntensity(rs=rs_ref)]
        draws = []
        for i in range(4):
            img = g.rand(*shape).astype(np.float32)
            sample = {"image": img.copy()}
            for t in chain:
                sample = t(sample)
            want = np.asarray(sample["image"])
            got, d = O.augment_sample(img, rs_mine, **kw)
            assert want.dtype == got.dtype == np.float32 and np.array_equal(want, got), (name, i)
            draws.append([d["top"], d["left"], d["gain"], d["cutoff"], d["k"]])
            out["%s_out%d" % (name, i)] = want[:, ::7, ::5]
        out["%s_draws" % name] = np.array(draws, dtype=np.float64)
        out["%s_shape" % name] = np.array(shape)
    out["input_seed"], out["rs_seed"] = np.array(21), np.array(4321)
    _save("augment_pins.npz", **out)


def gold_thick_slices():
    """LR-dataset synthesis: datasets/common_brains.py::simulate_thick_slices (scipy gaussian_filter1d per column)."""
    rb.bootstrap()
    from datasets.common_brains import simulate_thick_slices
    rs = np.random.RandomState(31)
    out = {}
    for k, (shape, thick) in enumerate((((24, 9, 11), 2.0), ((24, 9, 11), 3.0), ((7, 5, 6), 5.0), ((40, 6, 7), 6.0),
                                        ((3, 4, 5), 4.0))):
        vol = rs.rand(*shape).astype(np.float32)
        want = simulate_thick_slices(vol, thick)
        got = O.simulate_thick_slices(vol, thick)
        assert want.dtype == got.dtype == np.float32 and np.array_equal(want, got), (shape, thick)
        out["thick%d" % k] = want
        out["cfg%d" % k] = np.array(list(shape) + [thick], dtype=np.float64)
    out["seed"] = np.array(31)
    _save("thick_slices_pins.npz", **out)


def gold_sampling():
    """Pair / triplet sampling of the datasets' __getitem__ (RandomState draw order), run through the reference classes'
    own methods on a mock dataset object (no image files needed)."""
    rb.bootstrap()
    import types
    from datasets.common_brains import BrainDataset
    from datasets.ACDC.data4d_simple import ACDCDataset4DPairs
    out = {}
    for kind, sel, ds, Z in (("brain", "adjacent_plus", 4, 44), ("brain", "mix", 2, 30), ("brain", "adjacent_plus", 2, 9),
                             ("acdc", "adjacent_plus", 2, 10), ("acdc", "mix", 2, 8), ("acdc", "adjacent", 2, 6)):
        rs_ref, rs_mine = np.random.RandomState(2024), np.random.RandomState(2024)
        rows = []
        vol = np.broadcast_to(np.arange(Z, dtype=np.float32)[:, None, None], (Z, 2, 2)).copy()   # slice z holds the value z
        if kind == "brain":
            first = 1 if sel == "adjacent" else 0
            mock = types.SimpleNamespace(rs=rs_ref, slice_selection=sel, downsample_steps=ds, transform=None,
                                         images={0: {"image": vol}}, _idcs=[(0, z, Z) for z in range(Z)])
            mock._get_slice_step = lambda m=mock: BrainDataset._get_slice_step(m)
            mock._get_inbetween_sliceid = lambda a, b, m=mock: BrainDataset._get_inbetween_sliceid(m, a, b)
            cls = BrainDataset
        else:
            mock = types.SimpleNamespace(rs=rs_ref, slice_selection=sel, transform=None, _get_masks=False,
                                         images4d={0: {"image": vol[None], "orig_num_frames": 1, "num_slices": Z,
                                                       "spacing": None, "original_spacing": None, "patient_id": 0}},
                                         _idcs=np.array([(0, 0, z, Z) for z in range(Z)]))
            mock._get_slice_step = lambda m=mock: ACDCDataset4DPairs._get_slice_step(m)
            mock._get_inbetween_sliceid = ACDCDataset4DPairs._get_inbetween_sliceid
            cls = ACDCDataset4DPairs
        for rep in range(3):
            for z in range(Z):
                if kind == "brain" and (min(z + ds, Z - 1) - z < 2 and z - max(z - ds, 0) < 2) and sel != "mix":
                    continue
                try:
                    smp = cls.__getitem__(mock, z)
                except ValueError:                       # empty open interval (adjacent pair in a brain set): no triplet
                    rs_mine.set_state(rs_ref.get_state())
                    continue
                mine = O.sample_triplet(z, Z, rs_mine, kind=kind, slice_selection=sel, downsample_steps=ds)
                f = smp["slice_idx_from"] if kind == "brain" else int(smp["slice_id_from"][0])
                t = smp["slice_idx_to"] if kind == "brain" else int(smp["slice_id_to"][0])
                b = smp["inbetween_slice_id"] if kind == "brain" else None
                if kind == "acdc":       # the ACDC sample does not carry the in-between id: read it off the stacked slices
                    b = int(smp["image"][2, 0, 0])
                assert [int(v) for v in smp["image"][:, 0, 0]] == [int(f), int(t), int(b)]
                assert (int(f), int(t), int(b)) == (mine["slice_idx_from"], mine["slice_idx_to"], mine["inbetween_slice_id"]), (kind, sel, z)
                assert float(smp["is_inbetween"]) == float(mine["is_inbetween"])
                assert np.float32(smp["alpha_from"][0]) == mine["alpha_from"] and np.float32(smp["alpha_to"][0]) == mine["alpha_to"]
                rows.append([z, int(f), int(t), int(b), float(mine["is_inbetween"]), float(mine["alpha_from"]), float(mine["alpha_to"])])
        assert rs_ref.randint(0, 1 << 30) == rs_mine.randint(0, 1 << 30), (kind, sel)
        out["%s_%s_%d_%d" % (kind, sel, ds, Z)] = np.array(rows, dtype=np.float64)
    out["seed"] = np.array(2024)
    _save("sampling_pins.npz", **out)


def _sig(cls):
    return True


ALL = {"init": gold_init, "lpips": gold_lpips_lin, "infer": gold_infer, "train_small": gold_train_small,
       "transforms": gold_transforms, "augment": gold_augment, "thick": gold_thick_slices, "sampling": gold_sampling, "vif": gold_vif, "train_acdc": gold_train_acdc, "trained": gold_trained}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    if not rb.available():
        sys.exit("reference not available; fixtures are committed, nothing to do")
    rb.bootstrap()
    for name, fn in ALL.items():
        if a.only is not None and name not in a.only:
            continue
        if name in ("train_acdc", "trained") and (a.only is None or name not in a.only):
            continue                    # hundreds of CPU train steps (minutes): only on request
        print("== %s" % name, flush=True)
        fn(a.steps) if name == "train_acdc" else fn()
