"""Device-resident evaluation / data-path utilities: SSIM, PSNR, VIF and LPIPS per slice, the per-volume metric sets of
the model-selection loop, percentile normalisation, pad / crop.

Drop-ins for ``evaluate.metrics.compute_ssim_for_batch / compute_psnr_for_batch`` (evaluate/metrics.py:111-194),
``generate_hr_volumes.normalize_img`` (:130-133) / ``datasets.common.rescale_intensities`` (:408-417) and the crop / pad
transforms of ``datasets/shared_transforms.py`` (AdjustToPatchSize :389-447, CenterCrop :297-363, RandomCrop :48-120).
The reference does these on the host (numpy / scikit-image, one Python call per slice); here each is one pass of a
coalesced kernel over data that already sits in HBM after synthesis.

SSIM caveat (parity unpinned): the reference calls scikit-image without ``data_range`` on float
images, i.e. the legacy dtype range 2.0; ``data_range`` defaults to that and 1.0 can be passed for the images' true
range.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .ops import _dev, _stream


def _as_dev_f32(x, device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


def original_slice_ids(num_slices: int, downsample_steps: int, conv_interpol: bool = False) -> np.ndarray:
    """evaluate/metrics.py:29-45 (host index bookkeeping, bit-exact integer logic)."""
    ids = np.arange(num_slices)
    keep = None
    if (num_slices - 1) % downsample_steps != 0:
        r = (num_slices - 1) % downsample_steps
        keep, ids = ids[-r:], ids[:-r]
    if conv_interpol and ids.shape[0] % downsample_steps != 0:
        r = ids.shape[0] % downsample_steps
        keep = ids[-r:] if keep is None else np.concatenate((ids[-r:], keep))
        ids = ids[:-r]
    ids = ids[::downsample_steps]
    return ids if keep is None else np.concatenate((ids, keep))


def synth_slices_mask(orig_num_slices: int, downsample_steps: int):
    """evaluate/quantitative_comparison.py:10-17: (reconstructed-slice mask, synthesized-slice mask)."""
    n = ((orig_num_slices - 1) // downsample_steps) * downsample_steps + 1
    s_mask = np.ones(n, dtype=bool)
    s_mask[::downsample_steps] = False
    return ~s_mask, s_mask


def ssim_psnr_slices(true_vol, test_vol, win: int = 7, data_range: float = 2.0, device="cuda:0"):
    """Per-slice SSIM and PSNR of two [Z,H,W] volumes -> (ssim[Z], psnr[Z]) float64 numpy arrays."""
    a = _as_dev_f32(true_vol, device).squeeze()
    b = _as_dev_f32(test_vol, device).squeeze()
    if a.dim() == 2:
        a, b = a[None], b[None]
    assert a.shape == b.shape and a.dim() == 3
    lib = _dev(a)
    z, h, w = a.shape
    ssim_sum = torch.empty(z, dtype=torch.float64, device=a.device)
    sq = torch.empty(z, dtype=torch.float64, device=a.device)
    mk = torch.empty(z, dtype=torch.int32, device=a.device)
    _lib.check(lib.aesr_ssim_psnr(a.data_ptr(), b.data_ptr(), z, h, w, win, float(data_range), ssim_sum.data_ptr(),
                                  sq.data_ptr(), mk.data_ptr(), _stream(a)), "ssim_psnr")
    pad = (win - 1) // 2
    ssim = ssim_sum.cpu().numpy() / ((h - 2 * pad) * (w - 2 * pad))
    mse = sq.cpu().numpy() / (h * w)
    min_nonneg = mk.cpu().numpy().view(np.uint32) >= np.uint32(0x80000000)     # order-preserving key of min(true)
    rng = np.where(min_nonneg, 1.0, 2.0)               # skimage: float images, data_range 1 if min >= 0 else 2
    with np.errstate(divide="ignore"):
        psnr = 10 * np.log10(rng * rng / mse)
    return ssim, psnr


def compute_ssim_for_batch(l_images, l_reconstructions, eval_axis=0, normalize=False, downsample_steps=None,
                           conv_interpol=False, data_range: float = 2.0, device="cuda:0"):
    """evaluate/metrics.py:111-156 (eval_axis=0): mean SSIM over the non-original slices."""
    if eval_axis != 0 or normalize:
        raise NotImplementedError("aesr_b200: eval_axis != 0 / normalize=True are long-axis evaluation options "
                                  "outside the hot path")
    ssim, _ = ssim_psnr_slices(l_images, l_reconstructions, data_range=data_range, device=device)
    skip = original_slice_ids(len(ssim), downsample_steps, conv_interpol) if downsample_steps else []
    keep = np.setdiff1d(np.arange(len(ssim)), skip)
    return float(np.mean(ssim[keep]))


def compute_psnr_for_batch(l_images, l_reconstructions, eval_axis=0, normalize=False, downsample_steps=None,
                           conv_interpol=False, device="cuda:0"):
    """evaluate/metrics.py:159-194 (eval_axis=0): mean PSNR over the non-original slices, nan / inf dropped."""
    if eval_axis != 0 or normalize:
        raise NotImplementedError("aesr_b200: eval_axis != 0 / normalize=True are outside the hot path")
    _, psnr = ssim_psnr_slices(l_images, l_reconstructions, device=device)
    skip = original_slice_ids(len(psnr), downsample_steps, conv_interpol) if downsample_steps else []
    keep = np.setdiff1d(np.arange(len(psnr)), skip)
    vals = psnr[keep]
    vals = vals[np.isfinite(vals)]
    return float(np.mean(vals))


# ------------------------------------------------------------------------------------------------------------------
# VIF (evaluate/vifvec.py:7-63 through evaluate/metrics.py:65-108) and LPIPS as a metric (evaluate/metrics.py:210-242)
# ------------------------------------------------------------------------------------------------------------------
VIF_SIGMA_NSQ = 2.0
_VIF_FILTERS = {}


def _vif_filters(device):
    """The four gaussian kernels of vifp_mscale (N = 17, 9, 5, 3 -> sd = N/5, radius int(4 sd + .5)), formed on the host
    exactly like scipy.ndimage._filters._gaussian_kernel1d, concatenated on the device."""
    key = str(device)
    if key not in _VIF_FILTERS:
        ws, radii = [], []
        for scale in range(1, 5):
            sd = (2 ** (4 - scale + 1) + 1) / 5.0
            lw = int(4.0 * float(sd) + 0.5)
            x = np.arange(-lw, lw + 1)
            phi = np.exp(-0.5 / (sd * sd) * x ** 2)
            ws.append(phi / phi.sum())
            radii.append(lw)
        _VIF_FILTERS[key] = (torch.from_numpy(np.concatenate(ws)).to(device), np.asarray(radii, dtype=np.int32))
    return _VIF_FILTERS[key]


def vif_slices(true_vol, test_vol, device="cuda:0") -> np.ndarray:
    """Per-slice pixel-domain multi-scale VIF of two fp32 [Z,H,W] volumes in [0,1], exactly as the reference computes it
    on np.uint8(np.clip(x * 255, 0, 255)) slices -> float64 [Z] (nan where the denominator is 0, e.g. black slices)."""
    a = _as_dev_f32(true_vol, device).squeeze()
    b = _as_dev_f32(test_vol, device).squeeze()
    if a.dim() == 2:
        a, b = a[None], b[None]
    assert a.shape == b.shape and a.dim() == 3
    lib = _dev(a)
    z, h, w = a.shape
    st = _stream(a)
    a8 = torch.empty((z, h, w), dtype=torch.uint8, device=a.device)
    b8 = torch.empty((z, h, w), dtype=torch.uint8, device=a.device)
    _lib.check(lib.aesr_vif_quantize_u8(a.data_ptr(), a8.data_ptr(), a.numel(), st), "vif_quantize_u8")
    _lib.check(lib.aesr_vif_quantize_u8(b.data_ptr(), b8.data_ptr(), b.numel(), st), "vif_quantize_u8")
    wts, radii = _vif_filters(a.device)
    ws = torch.empty(int(lib.aesr_vif_workspace_bytes(z, h, w)), dtype=torch.uint8, device=a.device)
    nd = torch.empty((z, 2), dtype=torch.float64, device=a.device)
    _lib.check(lib.aesr_vif_mscale(a8.data_ptr(), b8.data_ptr(), z, h, w, wts.data_ptr(), radii.ctypes.data,
                                   float(VIF_SIGMA_NSQ), ws.data_ptr(), ws.numel(), nd.data_ptr(), st), "vif_mscale")
    nd = nd.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(nd[:, 1] != 0, nd[:, 0] / nd[:, 1], np.nan)


def compute_vif_for_batch(l_images, l_reconstructions, eval_axis=0, normalize=False, downsample_steps=None,
                          conv_interpol=False, device="cuda:0"):
    """evaluate/metrics.py:65-108 (eval_axis=0): mean VIF over the non-original slices, nan / inf dropped."""
    if eval_axis != 0 or normalize:
        raise NotImplementedError("aesr_b200: eval_axis != 0 / normalize=True are outside the hot path")
    vif = vif_slices(l_images, l_reconstructions, device=device)
    skip = original_slice_ids(len(vif), downsample_steps, conv_interpol) if downsample_steps else []
    vals = vif[np.setdiff1d(np.arange(len(vif)), skip)]
    vals = vals[np.isfinite(vals)]
    with np.errstate(invalid="ignore"):
        return float(np.mean(vals)) if vals.size else float("nan")


def compute_lpips_for_batch(l_images, l_reconstructions, eval_axis=0, normalize=False, downsample_steps=None,
                            conv_interpol=False, criterion=None, device="cuda:0"):
    """evaluate/metrics.py:210-242 (eval_axis=0): mean over the non-original slices of
    ``criterion(image_slice, recon_slice, normalize=True)`` -- here ONE batched pass of the sm_100a LPIPS-VGG forward
    (the reference loops over slices; the distance is per image, so batching cannot change a value)."""
    if eval_axis != 0 or normalize:
        raise NotImplementedError("aesr_b200: eval_axis != 0 / normalize=True are outside the hot path")
    if criterion is None:
        from .lpips_b200 import PerceptualLoss
        criterion = PerceptualLoss(model="net-lin", net="vgg", device=device)
    a = _as_dev_f32(l_images, device).squeeze()
    b = _as_dev_f32(l_reconstructions, device).squeeze()
    if a.dim() == 2:
        a, b = a[None], b[None]
    skip = original_slice_ids(a.shape[0], downsample_steps, conv_interpol) if downsample_steps else []
    keep = torch.from_numpy(np.setdiff1d(np.arange(a.shape[0]), skip)).to(a.device)
    vals = criterion(a[keep][:, None].contiguous(), b[keep][:, None].contiguous(), normalize=True)
    return float(np.mean(vals.reshape(-1).cpu().numpy().astype(np.float64)))


def compute_mean_metrics(ssim_results, psnr_results, vif_results, lpips_results, compute_percept_loss=False):
    """evaluate/create_HR_images.py:110-118."""
    out = []
    for r in (ssim_results, psnr_results, vif_results):
        out += [np.mean(np.array(r)), np.std(np.array(r))]
    if compute_percept_loss:
        out += [np.mean(np.array(lpips_results)), np.std(np.array(lpips_results))]
    else:
        out += [0, 0]
    return tuple(out)


def compute_metrics(images_ref, new_images, downsample_steps, ssim_results, psnr_results, vif_results, lpips_results,
                    ssim_res_synth=None, psnr_res_synth=None, vif_res_synth=None, lpips_res_synth=None,
                    ssim_res_recon=None, psnr_res_recon=None, vif_res_recon=None, lpips_res_recon=None,
                    compute_percept_loss=False, percept_loss=None, normalize=False, eval_axis=0, device="cuda:0"):
    """evaluate/create_HR_images.py:121-178, same signature and list side effects.  Per volume: SSIM / PSNR / VIF (and
    LPIPS on request) over all slices up to the last synthesised pair, over the synthesised slices only and over the
    reconstructed ones only (the masked LPIPS lists stay empty exactly like the reference, :172,177).  The per-slice
    values of SSIM, PSNR and VIF are computed ONCE on the device for the whole volume; the three slice sets are masks
    over them (each metric is per slice, so the subsets see the same numbers as three separate reference calls)."""
    if eval_axis != 0 or normalize:
        raise NotImplementedError("aesr_b200: eval_axis != 0 / normalize=True are outside the hot path")
    a = _as_dev_f32(images_ref, device).squeeze()
    b = _as_dev_f32(new_images, device).squeeze()
    last = ((a.shape[0] - 1) // downsample_steps) * downsample_steps + 1
    r_mask, s_mask = synth_slices_mask(a.shape[0], downsample_steps)
    a, b = a[:last].contiguous(), b[:last].contiguous()
    ssim, psnr = ssim_psnr_slices(a, b, device=device)
    vif = vif_slices(a, b, device=device)

    def finite_mean(v):
        v = v[np.isfinite(v)]
        with np.errstate(invalid="ignore"):
            return float(np.mean(v)) if v.size else float("nan")

    def push(mask, l_ssim, l_psnr, l_vif):
        l_ssim.append(float(np.mean(ssim[mask])))
        l_psnr.append(finite_mean(psnr[mask]))
        l_vif.append(finite_mean(vif[mask]))

    push(np.ones(last, dtype=bool), ssim_results, psnr_results, vif_results)
    if compute_percept_loss:
        lpips_results.append(compute_lpips_for_batch(a, b, criterion=percept_loss, device=device))
    if ssim_res_synth is not None:
        push(s_mask, ssim_res_synth, psnr_res_synth, vif_res_synth)
    if ssim_res_recon is not None:
        push(r_mask, ssim_res_recon, psnr_res_recon, vif_res_recon)
    return (ssim_results, psnr_results, vif_results, lpips_results, ssim_res_synth, psnr_res_synth, vif_res_synth,
            lpips_res_synth, ssim_res_recon, psnr_res_recon, vif_res_recon, lpips_res_recon)


def normalize_img(img, perc=(1, 99), device="cuda:0", return_percentiles: bool = False):
    """generate_hr_volumes.py:130-133 on the device: exact np.percentile (linear) over the whole volume + clip."""
    x = _as_dev_f32(img, device)
    lib = _dev(x)
    ws = torch.empty(int(lib.aesr_percentile_workspace_bytes()), dtype=torch.uint8, device=x.device)
    out = torch.empty_like(x)
    lo_hi = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(lib.aesr_percentile_normalize(x.data_ptr(), out.data_ptr(), x.numel(), float(perc[0]), float(perc[1]),
                                             ws.data_ptr(), ws.numel(), lo_hi.data_ptr(), _stream(x)),
               "percentile_normalize")
    return (out, lo_hi) if return_percentiles else out


def rescale_intensities(im, percs=(0, 100), device="cuda:0"):
    """datasets/common.py:408-417."""
    return normalize_img(im, percs, device=device)


def pad_crop(images, top, left, out_h: int, out_w: int, device="cuda:0") -> torch.Tensor:
    """out[b,c,y,x] = images[b,c,y+top[b],x+left[b]] or 0 outside; images [B,C,H,W] fp32."""
    x = _as_dev_f32(images, device)
    lib = _dev(x)
    b, c, h, w = x.shape
    t = torch.as_tensor(np.broadcast_to(np.asarray(top, dtype=np.int32), (b,)).copy(), device=x.device)
    l_ = torch.as_tensor(np.broadcast_to(np.asarray(left, dtype=np.int32), (b,)).copy(), device=x.device)
    out = torch.empty((b, c, out_h, out_w), dtype=torch.float32, device=x.device)
    _lib.check(lib.aesr_pad_crop_gather(x.data_ptr(), out.data_ptr(), t.data_ptr(), l_.data_ptr(), b, c, h, w, out_h,
                                        out_w, _stream(x)), "pad_crop_gather")
    return out


def adjust_to_patch_size(images, patch: int, device="cuda:0") -> torch.Tensor:
    """AdjustToPatchSize (shared_transforms.py:389-447): zero-pad up to >= patch, left = floor(d/2), right = ceil."""
    h, w = images.shape[-2:]
    dh, dw = max(patch - h, 0), max(patch - w, 0)
    return pad_crop(images, -(dh // 2), -(dw // 2), h + dh, w + dw, device=device)


def center_crop(images, patch: int, device="cuda:0") -> torch.Tensor:
    """CenterCrop (shared_transforms.py:297-363): window int(h/2) +- int(P/2)."""
    h, w = images.shape[-2:]
    half = int(patch / 2)
    return pad_crop(images, int(h / 2) - half, int(w / 2) - half, 2 * half, 2 * half, device=device)


def random_crop(images, patch: int, rs: np.random.RandomState, device="cuda:0") -> torch.Tensor:
    """RandomCrop (shared_transforms.py:48-120): one (top, left) per sample from rs.randint(0, h - P) (exclusive),
    same window for all channels of the sample (the from / to / between triplet)."""
    b, _, h, w = images.shape
    if h == patch and w == patch:
        return _as_dev_f32(images, device)
    tops, lefts = [], []
    for _ in range(b):
        tops.append(rs.randint(0, h - patch))
        lefts.append(rs.randint(0, w - patch))
    return pad_crop(images, np.array(tops), np.array(lefts), patch, patch, device=device)


def augment_window(h: int, w: int, aug_patch: Optional[int], center: bool):
    """Composite window of AdjustToPatchSize(aug) [-> CenterCrop(aug)]: (offset_y, offset_x, height, width) of the region the
    random crop is drawn from, in the coordinates of the unpadded h x w sample (zero padding: left = floor(d/2))."""
    off_y = off_x = 0
    hh, ww = h, w
    if aug_patch is not None:
        dh, dw = max(aug_patch - h, 0), max(aug_patch - w, 0)
        off_y, off_x, hh, ww = -(dh // 2), -(dw // 2), h + dh, w + dw
        if center:
            half = int(aug_patch / 2)
            off_y, off_x = off_y + int(hh / 2) - half, off_x + int(ww / 2) - half
            hh = ww = 2 * half
    return off_y, off_x, hh, ww


def augment_draw(rs: np.random.RandomState, hh: int, ww: int, width: int, intensity_first: bool):
    """One sample's draws in the order the reference's transform objects take them: RandomCrop randint(0,h-P),
    randint(0,w-P) (none when the sample already has the crop size); RandomIntensity uniform(2.5,7.5), uniform(.25,.75);
    RandomRotation randint(0,4) -- intensity before rotation for ACDC, after it for the brain sets.
    Returns (top, left, k, gain, cutoff) relative to the window of ``augment_window``."""
    if hh == width and ww == width:
        t = l_ = 0
    else:
        t = rs.randint(0, hh - width)
        l_ = rs.randint(0, ww - width)
    if intensity_first:
        g, cu = rs.uniform(2.5, 7.5), rs.uniform(0.25, 0.75)
        k = rs.randint(0, 4)
    else:
        k = rs.randint(0, 4)
        g, cu = rs.uniform(2.5, 7.5), rs.uniform(0.25, 0.75)
    return t, l_, k, g, cu


def augment_batch(images, rs: Optional[np.random.RandomState], width: int, aug_patch: Optional[int] = None,
                  center: bool = False, intensity_first: bool = True, slice_mask=None, device="cuda:0",
                  return_draws: bool = False, draws=None):
    """The training transform chain of the reference on a batch [B,C,H,W] of equally sized samples, one kernel:
    ACDC (train_cardiac_aesr.py:90-96): AdjustToPatchSize(aug) -> CenterCrop(aug) -> RandomCrop(width) -> RandomIntensity
    -> RandomRotation (``intensity_first=True, center=True``); brains (datasets/common_brains.py:55-57,77-80):
    [AdjustToPatchSize(aug)] -> RandomCrop(width) -> RandomRotation -> RandomIntensity.  The draws are taken from ``rs``
    on the host, sample by sample, in exactly the order the reference's transform objects take them (``augment_draw``),
    so a seeded RandomState yields the reference's augmentation stream; a caller that interleaves them with other draws
    (``data_loader.DeviceTripletLoader``: triplet indices, then transforms, per sample like ``__getitem__``) passes them as
    ``draws`` = B tuples of ``augment_draw``.  ``slice_mask``: boolean per channel (ACDCLBL).  Returns fp32
    [B,C,width,width] on ``device`` (and the draws with ``return_draws``)."""
    x = _as_dev_f32(images, device)
    lib = _dev(x)
    b, c, h, w = x.shape
    off_y, off_x, hh, ww = augment_window(h, w, aug_patch, center)
    if draws is None:
        draws = [augment_draw(rs, hh, ww, width, intensity_first) for _ in range(b)]
    assert len(draws) == b
    tops = [d[0] + off_y for d in draws]
    lefts = [d[1] + off_x for d in draws]
    ks, gains, cuts = [d[2] for d in draws], [d[3] for d in draws], [d[4] for d in draws]
    mask = 0xFFFFFFFF
    if slice_mask is not None:
        sm = np.asarray(slice_mask, dtype=bool)
        assert sm.shape == (c,) and c <= 32
        mask = int(sum(1 << i for i in range(c) if sm[i]))
    ti = torch.as_tensor(np.asarray([tops, lefts, ks], dtype=np.int32), device=x.device)
    # python floats are "weak" scalars against float32 arrays in numpy: gain and cutoff act as float32
    tf = torch.as_tensor(np.asarray([gains, cuts], dtype=np.float64).astype(np.float32), device=x.device)
    out = torch.empty((b, c, width, width), dtype=torch.float32, device=x.device)
    _lib.check(lib.aesr_augment_gather(x.data_ptr(), out.data_ptr(), ti[0].data_ptr(), ti[1].data_ptr(), ti[2].data_ptr(),
                                       tf[0].data_ptr(), tf[1].data_ptr(), mask, b, c, h, w, width, _stream(x)),
               "augment_gather")
    if return_draws:
        return out, {"top": np.asarray(tops) - off_y, "left": np.asarray(lefts) - off_x, "gain": np.asarray(gains),
                     "cutoff": np.asarray(cuts), "k": np.asarray(ks)}
    return out


def prepare_batch_pairs(batch_dict: dict, expand_type: str = "repeat") -> dict:
    """datasets/common_brains.py:285-321 / datasets/ACDC/data4d_simple.py:327-387, same signature and in-place dict
    semantics: batch_dict['image'] [B,2|3,H,W] -> 'image' [2B,1,H,W] (all "from" slices, then all "to" slices) and
    'slice_between' [B,1,H,W]; 'split' keeps 'image' and adds 'image_from' / 'image_to'.  Works on device tensors (the
    output of ``augment_batch``): no host round trip between augmentation and the training step."""
    batch_images = batch_dict["image"]
    assert batch_images.size(0) % 2 == 0
    if expand_type not in ("repeat", "split"):
        raise ValueError("Error - prepare_batch_pairs - valid values for expand_type parameter are repeat, "
                         "reshape, split.")
    a = torch.unsqueeze(batch_images[:, 0], dim=1)
    b = torch.unsqueeze(batch_images[:, 1], dim=1)
    if batch_images.shape[1] == 3:
        batch_dict["slice_between"] = torch.unsqueeze(batch_images[:, 2], dim=1)
    if expand_type == "split":
        batch_dict["image_from"], batch_dict["image_to"] = a, b
    else:
        batch_dict["image"] = torch.cat([a, b], dim=0)
    return batch_dict


def determine_interpol_coefficients(sliceid_from, sliceid_to, sliceid_between):
    """datasets/common_brains.py:117-119 (float64; the brain trainers cast to float32, :259-260)."""
    gap = sliceid_to - sliceid_from
    return 1 - ((sliceid_between - sliceid_from) * 1 / gap), 1 - ((sliceid_to - sliceid_between) * 1 / gap)


def simulate_thick_slices(img3d, slice_thickness: float, device="cuda:0") -> torch.Tensor:
    """datasets/common_brains.py:37-44: thick-slice simulation of a volume [Z,H,W] -- a gaussian slice profile of
    FWHM = ``slice_thickness`` along z for every (y, x) column (scipy.ndimage.gaussian_filter1d semantics, bit-exact:
    'reflect' borders, float64 accumulation in scipy's order).  Returns fp32 [Z,H,W] on ``device``."""
    x = _as_dev_f32(img3d, device)
    lib = _dev(x)
    assert x.dim() == 3
    sd = slice_thickness / 2.355
    lw = int(4.0 * float(sd) + 0.5)
    k = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * k ** 2)             # scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius)
    taps = torch.as_tensor(phi / phi.sum(), dtype=torch.float64, device=x.device)
    out = torch.empty_like(x)
    z, h, w = x.shape
    _lib.check(lib.aesr_gauss1d_axis0(x.data_ptr(), out.data_ptr(), taps.data_ptr(), lw, z, h * w, _stream(x)),
               "gauss1d_axis0")
    return out
