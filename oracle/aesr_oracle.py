"""CPU oracle for the ae_combined slice-synthesis hot path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (torch-CPU fp32 / numpy float64) of the
reference's algorithm for the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker /
CPU baseline -- never as the product path.  Nothing in here touches ``/root/reference`` at run time.

Parity pin: the restatement is checked against outputs of the *unmodified reference code* imported in
the build container by ``oracle/make_golden.py`` (which also writes ``tests/golden/*.npz``).  The one
exception is SSIM/PSNR: the reference delegates to scikit-image, which is neither vendored nor pinned
nor installed here, so that part is **parity unpinned** (restated from the published algorithm, see
``ssim_slice`` below).

Every function cites the reference file:line it follows (paths relative to the reference repo root).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01      # nn.LeakyReLU() default, networks/acai_vanilla.py:17
BN_EPS = 1e-5           # nn.BatchNorm2d default, networks/acai_vanilla.py:58,90
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# architecture tables
# ----------------------------------------------------------------------------------------------
def num_scales(width: int, latent_width: int) -> int:
    """networks/acai_vanilla.py:116."""
    return int(round(math.log(width // latent_width, 2)))


def encoder_spec(scales: int, depth: int, latent: int, colors: int = 1) -> List[tuple]:
    """Layer list of ``Encoder`` (networks/acai_vanilla.py:49-72, use_batchnorm=True, n_res_block=None).

    Entries: ("conv", idx, cin, cout, ksize, pad, act) | ("bn", idx, c) | ("pool",) with ``idx`` the
    nn.Sequential position (= state_dict key)."""
    spec, idx = [], 0
    spec.append(("conv", idx, colors, depth, 1, 1, None)); idx += 1
    kp = depth
    for s in range(scales):
        k = depth << s
        spec.append(("conv", idx, kp, k, 3, 1, "leaky")); idx += 2
        spec.append(("conv", idx, k, k, 3, 1, "leaky")); idx += 2
        spec.append(("bn", idx, k)); idx += 1
        spec.append(("pool",)); idx += 1
        kp = k
    k = depth << scales
    spec.append(("conv", idx, kp, k, 3, 1, "leaky")); idx += 2
    spec.append(("conv", idx, k, latent, 3, 1, None)); idx += 1
    return spec


def decoder_spec(scales: int, depth: int, latent: int, colors: int = 1) -> List[tuple]:
    """Layer list of ``Decoder`` (networks/acai_vanilla.py:75-102, use_upsample, use_batchnorm, use_sigmoid)."""
    spec, idx = [], 0
    kp = latent
    for s in range(scales - 1, -1, -1):
        k = depth << s
        spec.append(("conv", idx, kp, k, 3, 1, "leaky")); idx += 2
        spec.append(("conv", idx, k, k, 3, 1, "leaky")); idx += 2
        spec.append(("bn", idx, k)); idx += 1
        spec.append(("up",)); idx += 1
        kp = k
    spec.append(("conv", idx, kp, depth, 3, 1, "leaky")); idx += 2
    spec.append(("conv", idx, depth, colors, 3, 1, "sigmoid")); idx += 2
    return spec


def init_state(args: dict, seed: Optional[int] = None) -> "OrderedDict[str, torch.Tensor]":
    """Random-init ``state_dict`` exactly as ``VanillaACAI(args)`` produces it.

    Follows networks/acai_vanilla.py:39-46 (``Initializer``: conv AND BatchNorm weights ~ N(0, std) with
    std = 1/sqrt(1.04 * prod(shape[:-1])), biases zero) and :49-72/:75-102 for the construction order,
    which fixes the RNG stream (each nn.Conv2d's default init draws first, then ``normal_`` per layer)."""
    if seed is not None:
        torch.manual_seed(seed)
    scales = num_scales(args["width"], args["latent_width"])
    state: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for prefix, spec in (("enc", encoder_spec(scales, args["depth"], args["latent"], args.get("colors", 1))),
                         ("dec", decoder_spec(scales, args["depth"], args["latent"], args.get("colors", 1)))):
        layers = []
        for ent in spec:
            if ent[0] == "conv":
                _, idx, cin, cout, ks, pad, _ = ent
                layers.append((idx, torch.nn.Conv2d(cin, cout, ks, padding=pad)))
            elif ent[0] == "bn":
                layers.append((ent[1], torch.nn.BatchNorm2d(ent[2])))
        for idx, layer in layers:           # Initializer pass, module order
            w = layer.weight.data
            std = 1 / np.sqrt((1 + 0.2 ** 2) * np.prod(w.shape[:-1]))
            w.normal_(std=std)
            layer.bias.data.zero_()
        for idx, layer in layers:
            for k, v in layer.state_dict().items():
                state["%s.%d.%s" % (prefix, idx, k)] = v.clone()
    return state


# ----------------------------------------------------------------------------------------------
# forward passes (functional, autograd-capable)
# ----------------------------------------------------------------------------------------------
def _run(spec, prefix, state, x, train: bool, bn_updates: Optional[dict], momentum: float = BN_MOMENTUM):
    for ent in spec:
        kind = ent[0]
        if kind == "conv":
            _, idx, _, _, _, pad, act = ent
            x = F.conv2d(x, state["%s.%d.weight" % (prefix, idx)], state["%s.%d.bias" % (prefix, idx)], padding=pad)
            if act == "leaky":
                x = F.leaky_relu(x, LEAKY_SLOPE)
            elif act == "sigmoid":
                x = torch.sigmoid(x)
        elif kind == "bn":
            key = "%s.%d" % (prefix, ent[1])
            rm, rv = state[key + ".running_mean"], state[key + ".running_var"]
            if train:
                # F.batch_norm updates rm/rv in place (momentum .1, unbiased var) like nn.BatchNorm2d.train()
                x = F.batch_norm(x, rm, rv, state[key + ".weight"], state[key + ".bias"], True, momentum, BN_EPS)
                state[key + ".num_batches_tracked"] += 1
            else:
                x = F.batch_norm(x, rm, rv, state[key + ".weight"], state[key + ".bias"], False, BN_MOMENTUM, BN_EPS)
        elif kind == "pool":
            x = F.avg_pool2d(x, 2)
        elif kind == "up":
            x = F.interpolate(x, scale_factor=2, mode="nearest")
    return x


def run_trace(spec, prefix, state, x: torch.Tensor) -> List[Tuple[str, torch.Tensor, torch.Tensor]]:
    """Eval-mode pass that records, for every conv of ``spec``, (state key prefix, conv INPUT, conv OUTPUT after its
    activation and the BatchNorm / pool / upsample entries that follow it up to the next conv) -- the per-layer error
    ledger of tests/test_gpu_parity.py feeds each of our fused layers the oracle's input and compares the outputs."""
    out, cur_in, cur_key = [], None, None
    for ent in spec:
        if ent[0] == "conv":
            if cur_key is not None:
                out.append((cur_key, cur_in, x))
            cur_key, cur_in = "%s.%d" % (prefix, ent[1]), x
        x = _run([ent], prefix, state, x, False, None)
    out.append((cur_key, cur_in, x))
    return out


def encode(state, args, x: torch.Tensor, train: bool = False) -> torch.Tensor:
    """VanillaACAI.encode, networks/acai_vanilla.py:134-135.  ``train=True`` mutates BN running stats in ``state``."""
    scales = num_scales(args["width"], args["latent_width"])
    return _run(encoder_spec(scales, args["depth"], args["latent"], args.get("colors", 1)), "enc", state, x, train, None)


def decode(state, args, z: torch.Tensor, train: bool = False) -> torch.Tensor:
    """VanillaACAI.decode, networks/acai_vanilla.py:137-138."""
    scales = num_scales(args["width"], args["latent_width"])
    return _run(decoder_spec(scales, args["depth"], args["latent"], args.get("colors", 1)), "dec", state, z, train, None)


# ----------------------------------------------------------------------------------------------
# volume synthesis (inference)
# ----------------------------------------------------------------------------------------------
def alpha_range_for(num_interpolations: int) -> np.ndarray:
    """generate_hr_volumes.py:162 (float64)."""
    return np.linspace(0, 1, num_interpolations + 2, endpoint=True)[1:-1]


def interp_weights(alpha_range: Sequence[float]) -> Tuple[np.ndarray, np.ndarray]:
    """Bit-exact fp32 lerp weights.  ``alpha * latent_1 + (1 - alpha) * latent_2`` with a python/numpy
    float64 ``alpha`` and fp32 tensors (generate_hr_volumes.py:88): torch converts each float64 scalar
    to fp32 once, so w_hi = fp32(alpha), w_lo = fp32(1 - alpha) (the subtraction happens in float64)."""
    a = np.asarray(alpha_range, dtype=np.float64)
    return a.astype(np.float32), (1.0 - a).astype(np.float32)


def latent_space_interp(state, args, alpha, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """generate_hr_volumes.py:72-101 / kwatsch/img_interpolation.py:57-89 (eval mode, no_grad)."""
    with torch.no_grad():
        latent_1 = encode(state, args, img1.float(), train=False)
        latent_2 = encode(state, args, img2.float(), train=False)
        inter = alpha * latent_1 + (1 - alpha) * latent_2
        return decode(state, args, inter, train=False)


def create_super_volume(state, args, images: torch.Tensor, alpha_range, use_original: bool = False) -> torch.Tensor:
    """generate_hr_volumes.py:12-69 (labels=None).  Returns ``upsampled_image`` [(Z-1)(A+1)+1, H, W]."""
    if images.dim() == 3:
        images = images.unsqueeze(1)
    num_slices = images.shape[0]
    if not use_original:
        with torch.no_grad():          # trainer.predict -> model(x) in eval mode, kwatsch/base_trainer.py:216-246
            recon = decode(state, args, encode(state, args, images.float()))
    else:
        recon = images
    images2, images1 = images[1:], images[:-1]
    interp = None
    for alpha in alpha_range:
        img = latent_space_interp(state, args, alpha, images2, images1)
        interp = img if interp is None else torch.cat([interp, img], dim=1)
    vol = None
    for i in range(num_slices - 1):
        vol = torch.cat([recon[i], interp[i]]) if vol is None else torch.cat([vol, recon[i], interp[i]], dim=0)
    vol = torch.cat([vol, recon[num_slices - 1]])
    return torch.clamp(vol, min=0, max=1.)


def create_super_volume_eval(state, args, images: torch.Tensor, alpha_range=None, use_original: bool = False,
                             downsample_steps: Optional[int] = None, generate_inbetween_slices: bool = False) -> torch.Tensor:
    """evaluate/common.py:134-235 (labels=None, hierarchical=False).  ``images`` is [Z, H, W]."""
    if generate_inbetween_slices and downsample_steps is None:
        downsample_steps = int(len(alpha_range) + 1)
    orig_images, orig_num = None, images.shape[0]
    if downsample_steps is not None or generate_inbetween_slices:
        orig_images = images.clone()
        if (orig_num - 1) % downsample_steps != 0:
            images = images[:-((orig_num - 1) % downsample_steps)]
        images = images[::downsample_steps]
    if alpha_range is None:
        alpha_range = [0.25, 0.5, 0.75]
    vol = create_super_volume(state, args, images, alpha_range, use_original=use_original)
    # create_super_volume clamps at the end; the eval twin appends the trimmed tail BEFORE the clamp, which is
    # the same thing for the concatenation (clamp is elementwise).
    if generate_inbetween_slices and (orig_num - 1) % downsample_steps != 0:
        remain = (orig_num - 1) % downsample_steps
        vol = torch.cat([vol, torch.clamp(orig_images[-remain:].float(), 0, 1.)])
    return vol


def synth_slice_mask(orig_num_slices: int, downsample_steps: int) -> Tuple[np.ndarray, np.ndarray]:
    """evaluate/quantitative_comparison.py:10-17 + evaluate/common.py:36-39 (determine_last_slice)."""
    last = ((orig_num_slices - 1) // downsample_steps) * downsample_steps
    n = last + 1
    s_mask = np.ones(n, dtype=bool)
    r_mask = np.zeros(n, dtype=bool)
    s_mask[np.arange(0, n)[::downsample_steps]] = False
    r_mask[np.arange(0, n)[::downsample_steps]] = True
    return r_mask, s_mask


def determine_original_sliceids(num_slices: int, downsample_steps: int, conv_interpol: bool = False) -> np.ndarray:
    """evaluate/metrics.py:29-45."""
    slice_ids = np.arange(num_slices)
    keep = None
    if (num_slices - 1) % downsample_steps != 0:
        r = (num_slices - 1) % downsample_steps
        keep = slice_ids[-r:]
        slice_ids = slice_ids[:-r]
    if conv_interpol and slice_ids.shape[0] % downsample_steps != 0:
        r = slice_ids.shape[0] % downsample_steps
        keep = slice_ids[-r:] if keep is None else np.concatenate((slice_ids[-r:], keep))
        slice_ids = slice_ids[:-r]
    slice_ids = slice_ids[::downsample_steps]
    if keep is not None:
        slice_ids = np.concatenate((slice_ids, keep))
    return slice_ids


# ----------------------------------------------------------------------------------------------
# normalisation / transforms / batch layout
# ----------------------------------------------------------------------------------------------
def normalize_img(img: np.ndarray, perc=(1, 99)) -> np.ndarray:
    """generate_hr_volumes.py:130-133: float64 percentiles => float64 division under numpy>=2."""
    lo, hi = np.percentile(img, perc)
    return ((img.astype(img.dtype) - lo) / (hi - lo)).clip(0, 1)


def sitk_array_to_torch(np_img: np.ndarray) -> torch.Tensor:
    """generate_hr_volumes.py:104-111 minus the SimpleITK read."""
    np_img = np_img.astype(np.float32)
    if np_img.max() > 1 or np_img.min() < 0:
        np_img = normalize_img(np_img)
    return torch.from_numpy(np_img).float().unsqueeze(1)


def rescale_intensities(im: np.ndarray, percs=(0, 100)) -> np.ndarray:
    """datasets/common.py:408-417."""
    lo, hi = np.percentile(im, percs)
    if np.isnan(lo):
        lo = 0
    if np.isnan(hi):
        hi = 1
    return ((im.astype(np.float32) - lo) / (hi - lo)).clip(0, 1)


def adjust_to_patch_size(img: np.ndarray, patch: int) -> np.ndarray:
    """datasets/shared_transforms.py:389-447 (AdjustToPatchSize): zero-pad [C,h,w] (or [h,w]) up to >= patch,
    left = floor(delta/2), right = ceil(delta/2)."""
    h, w = img.shape[-2:]
    dh, dw = max(patch - h, 0), max(patch - w, 0)
    pads = [(0, 0)] * (img.ndim - 2) + [(dh // 2, dh - dh // 2), (dw // 2, dw - dw // 2)]
    return np.pad(img, pads, mode="constant", constant_values=0)


def center_crop(img: np.ndarray, patch: int) -> np.ndarray:
    """datasets/shared_transforms.py:297-363 (CenterCrop): window int(w/2) +- int(P/2)."""
    h, w = img.shape[-2:]
    half = int(patch / 2)
    return img[..., int(h / 2) - half:int(h / 2) + half, int(w / 2) - half:int(w / 2) + half]


def random_crop_offsets(rs: np.random.RandomState, h: int, w: int, patch: int) -> Tuple[int, int]:
    """datasets/shared_transforms.py:48-120 (RandomCrop): rs.randint(0, h - P) (exclusive upper bound)."""
    if h == patch and w == patch:        # "rare case": sample returned untouched, no RNG draw
        return 0, 0
    top = rs.randint(0, h - patch)       # raises ValueError when h == patch != w, exactly like the reference
    left = rs.randint(0, w - patch)
    return top, left


def random_intensity(img: np.ndarray, gain: float, cutoff: float) -> np.ndarray:
    """datasets/shared_transforms.py:366-386 (RandomIntensity): sigmoid contrast 1/(1+exp(g*(c-x)))."""
    return 1.0 / (1.0 + np.exp(gain * (cutoff - img)))


def random_rotation(img: np.ndarray, k: int) -> np.ndarray:
    """datasets/shared_transforms.py:224-254 (RandomRotation): np.rot90 by k quarter turns in the last two axes."""
    return np.rot90(img, k, (img.ndim - 2, img.ndim - 1)).copy()


def augment_sample(img: np.ndarray, rs: np.random.RandomState, width: int, aug_patch: Optional[int] = None,
                   center: bool = False, intensity_first: bool = True, slice_mask=None) -> Tuple[np.ndarray, dict]:
    """The training transform chain of one sample [C,H,W] with the reference's RandomState draw order:
    ACDC (train_cardiac_aesr.py:90-96): AdjustToPatchSize(aug) -> CenterCrop(aug) -> RandomCrop(width) ->
    RandomIntensity -> RandomRotation (``intensity_first``); brains (datasets/common_brains.py:55-57,77-80):
    [AdjustToPatchSize(aug)] -> RandomCrop(width) -> RandomRotation -> RandomIntensity.
    Returns the augmented sample and the draws {top, left, gain, cutoff, k}."""
    x = img
    if aug_patch is not None:
        x = adjust_to_patch_size(x, aug_patch)
        if center:
            x = center_crop(x, aug_patch)
    top, left = random_crop_offsets(rs, x.shape[-2], x.shape[-1], width)
    x = x[..., top:top + width, left:left + width] if x.shape[-1] != width or x.shape[-2] != width else x

    def intensity(v):
        gain = rs.uniform(2.5, 7.5)
        cutoff = rs.uniform(0.25, 0.75)
        if slice_mask is None:
            return random_intensity(v, gain, cutoff), gain, cutoff
        v = v.copy()
        v[slice_mask] = random_intensity(v[slice_mask], gain, cutoff)
        return v, gain, cutoff

    if intensity_first:
        x, gain, cutoff = intensity(x)
        k = rs.randint(0, 4)
        x = random_rotation(x, k)
    else:
        k = rs.randint(0, 4)
        x = random_rotation(x, k)
        x, gain, cutoff = intensity(x)
    return x, {"top": top, "left": left, "gain": gain, "cutoff": cutoff, "k": k}


def get_random_adjacent_slice(slice_id: int, num_slices: int, rs: np.random.RandomState, step: int = 1):
    """datasets/common.py:34-43: the partner slice `step` away; a draw (rs.choice of two) only when both sides exist."""
    last = num_slices - 1
    if slice_id + step > last:
        return slice_id - step
    if slice_id == 0:
        return step
    if slice_id - step < 0:
        return slice_id + step
    return rs.choice([slice_id - step, slice_id + step])


def sample_triplet(slice_id_1: int, num_slices: int, rs: np.random.RandomState, kind: str = "acdc",
                   slice_selection: str = "adjacent_plus", downsample_steps: int = 2) -> dict:
    """One training triplet (from, to, between) with the reference's RandomState draw order:
    ACDC  datasets/ACDC/data4d_simple.py:191-205,245-263 (step 1 / 2 / choice([1,2]); midpoint or slice 1 itself, alphas 0.5)
    brain datasets/common_brains.py:241-260,272-282 (step 1 / downsample_steps / choice; in-between drawn from the open
    interval; alphas from determine_interpol_coefficients, cast to float32)."""
    if slice_selection == "adjacent":
        step = 1
    elif slice_selection == "adjacent_plus":
        step = 2 if kind == "acdc" else downsample_steps
    else:
        step = rs.choice([1, 2 if kind == "acdc" else downsample_steps])
    s2 = get_random_adjacent_slice(slice_id_1, num_slices, rs, step)
    if kind == "acdc":
        if (slice_id_1 + s2) % 2 == 0:
            between, is_between = (slice_id_1 + s2) // 2, 1
        else:
            between, is_between = slice_id_1, 0
    else:
        between, is_between = rs.choice(np.arange(min(slice_id_1, s2) + 1, max(slice_id_1, s2))), 1
    s_from, s_to = (slice_id_1, s2) if rs.choice([0, 1]) == 0 else (s2, slice_id_1)
    if kind == "acdc":
        a_from = a_to = np.float32(0.5)
    else:
        af, at = determine_interpol_coefficients(s_from, s_to, between)
        a_from, a_to = np.float32(af), np.float32(at)
    return {"slice_idx_from": int(s_from), "slice_idx_to": int(s_to), "inbetween_slice_id": int(between),
            "is_inbetween": is_between, "alpha_from": a_from, "alpha_to": a_to}


def prepare_batch_pairs(batch_images: torch.Tensor) -> Dict[str, torch.Tensor]:
    """datasets/common_brains.py:285-321 / datasets/ACDC/data4d_simple.py:327-387 ('repeat'):
    [B,3,H,W] -> image [2B,1,H,W] (all 'from' then all 'to'), slice_between [B,1,H,W]."""
    a = batch_images[:, 0:1]
    b = batch_images[:, 1:2]
    out = {"image": torch.cat([a, b], dim=0)}
    if batch_images.shape[1] == 3:
        out["slice_between"] = batch_images[:, 2:3]
    return out


def determine_interpol_coefficients(sliceid_from, sliceid_to, sliceid_between):
    """datasets/common_brains.py:117-119 (float64; callers cast to float32, :259-260)."""
    gap = sliceid_to - sliceid_from
    return 1 - ((sliceid_between - sliceid_from) * 1 / gap), 1 - ((sliceid_to - sliceid_between) * 1 / gap)


# ----------------------------------------------------------------------------------------------
# LPIPS-VGG v0.1 (net-lin)
# ----------------------------------------------------------------------------------------------
VGG_CFG = [(3, 64), (64, 64), "M", (64, 128), (128, 128), "M", (128, 256), (256, 256), (256, 256), "M",
           (256, 512), (512, 512), (512, 512), "M", (512, 512), (512, 512), (512, 512)]
VGG_TAPS_AFTER_CONV = (1, 3, 6, 9, 12)      # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (lpips/pretrained_networks.py:107-135)
LPIPS_SHIFT = (-.030, -.088, -.188)         # lpips/networks_basic.py:96
LPIPS_SCALE = (.458, .448, .450)            # lpips/networks_basic.py:97
LPIPS_CHNS = (64, 128, 256, 512, 512)


def init_vgg(seed: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Random-init VGG16 trunk as ``torchvision.models.vgg16(weights=None).features`` initialises it
    (kaiming_normal_(fan_out, relu), bias 0) restricted to the 13 convs LPIPS uses.  Drawn in module order
    from ``torch.manual_seed(seed)``; pinned against torchvision by oracle/make_golden.py."""
    torch.manual_seed(seed)
    convs = []
    for ent in VGG_CFG:
        if ent == "M":
            continue
        cin, cout = ent
        convs.append(torch.nn.Conv2d(cin, cout, 3, padding=1))      # default init draws from the RNG
    # torchvision builds the (unused) classifier before its init loop; its default inits advance the RNG too.
    for fin, fout in ((512 * 7 * 7, 4096), (4096, 4096), (4096, 1000)):
        torch.nn.Linear(fin, fout)
    for conv in convs:
        torch.nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
        torch.nn.init.constant_(conv.bias, 0)
    return [(c.weight.data.clone(), c.bias.data.clone()) for c in convs]


def vgg_taps(vgg, x3: torch.Tensor) -> List[torch.Tensor]:
    """lpips/pretrained_networks.py:121-135."""
    taps, ci, h = [], 0, x3
    for ent in VGG_CFG:
        if ent == "M":
            h = F.max_pool2d(h, 2)
            continue
        w, b = vgg[ci]
        h = F.relu(F.conv2d(h, w, b, padding=1))
        if ci in VGG_TAPS_AFTER_CONV:
            taps.append(h)
        ci += 1
    return taps


def lpips_forward(vgg, lins: Sequence[torch.Tensor], pred: torch.Tensor, target: torch.Tensor,
                  normalize: bool = True) -> torch.Tensor:
    """PerceptualLoss.forward (lpips/perceptual.py:19-33) -> PNetLin.forward (lpips/networks_basic.py:63-91).
    ``lins[k]`` is the [1,C_k,1,1] NetLinLayer weight (dropout inactive in eval).  Returns [N,1,1,1]."""
    if normalize:
        target = 2 * target - 1
        pred = 2 * pred - 1
    in0, in1 = target, pred                      # self.model.forward(target, pred)
    shift = torch.tensor(LPIPS_SHIFT, device=pred.device)[None, :, None, None]
    scale = torch.tensor(LPIPS_SCALE, device=pred.device)[None, :, None, None]
    o0 = vgg_taps(vgg, (in0 - shift) / scale)   # 1 -> 3 channel broadcast, lpips/networks_basic.py:99-100
    o1 = vgg_taps(vgg, (in1 - shift) / scale)
    val = None
    for k in range(5):
        f0 = o0[k] / (torch.sqrt(torch.sum(o0[k] ** 2, dim=1, keepdim=True)) + 1e-10)   # lpips/common.py:12-14
        f1 = o1[k] / (torch.sqrt(torch.sum(o1[k] ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (f0 - f1) ** 2
        r = F.conv2d(d, lins[k]).mean([2, 3], keepdim=True)
        val = r if val is None else val + r
    return val


# ----------------------------------------------------------------------------------------------
# training step
# ----------------------------------------------------------------------------------------------
PARAM_SUFFIXES = (".weight", ".bias")


def param_keys(state) -> List[str]:
    """``model.parameters()`` order = module order (enc then dec), weight then bias."""
    return [k for k in state.keys() if k.endswith(PARAM_SUFFIXES)]


class AdamState:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay) restated (kwatsch/trainer_ae.py:29-30)."""

    def __init__(self, state, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.wd, self.betas, self.eps, self.t = lr, weight_decay, betas, eps, 0
        self.m = {k: torch.zeros_like(state[k]) for k in param_keys(state)}
        self.v = {k: torch.zeros_like(state[k]) for k in param_keys(state)}

    def step(self, state, grads: Dict[str, torch.Tensor]):
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for k, g in grads.items():
            if self.wd != 0:
                g = g + self.wd * state[k]
            self.m[k].lerp_(g, 1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            state[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def train_step(state, args, adam: Optional[AdamState], image: torch.Tensor, slice_between: torch.Tensor,
               vgg=None, lins=None, alpha_from: Optional[torch.Tensor] = None, alpha_to: Optional[torch.Tensor] = None,
               ex_loss_weight: float = 0.05, combined: bool = True, return_grads: bool = False) -> dict:
    """One optimisation step.

    ``combined=True``  : AETrainerEndToEnd.train (kwatsch/cardiac/trainer_ae.py:10-50) when alpha_from is None
                         (z_mix = 0.5 z[:B] + (1-0.5) z[B:], :173), AETrainerExtension1Brain.train
                         (kwatsch/brain/trainer_ae.py:92-132, :265-266) when per-sample alphas [B,1] are given.
                         Order of train-mode passes (each updates BN running stats): enc(x) -> dec(z) ->
                         dec(z_mix) -> enc(slice_between).  loss = MSE(out,x) + w * mean(LPIPS(between, synth)).
    ``combined=False`` : AEBaseTrainer.train (kwatsch/trainer_ae.py:71-109): loss = MSE only; the logged latent loss
                         encodes slice_between in eval mode.
    Mutates ``state`` (params + BN buffers) and ``adam``.  Returns the logged scalars (+ grads if asked)."""
    keys = param_keys(state)
    params = {k: state[k].detach().clone().requires_grad_(True) for k in keys}
    work = OrderedDict((k, params.get(k, v)) for k, v in state.items())
    B = image.shape[0] // 2
    z = encode(work, args, image, train=True)
    out = decode(work, args, z, train=True)
    loss_dist = F.mse_loss(out, image, reduction="mean")            # kwatsch/base_trainer.py:177
    logs = {"loss_ae_dist": float(loss_dist.detach())}
    if combined:
        if alpha_from is None:
            a05 = torch.tensor([0.5], device=z.device)[:, None, None, None]
            z_mix = a05 * z[:B] + (1 - a05) * z[B:]
        else:
            z_mix = alpha_from[:, :, None, None] * z[:B] + alpha_to[:, :, None, None] * z[B:]
        s_mix = decode(work, args, z_mix, train=True)
        z_ref = encode(work, args, slice_between, train=True)       # logged only, but updates BN stats
        logs["loss_latent_1"] = float(F.mse_loss(z_mix, z_ref))
        lp = lpips_forward(vgg, lins, s_mix, slice_between, normalize=True).mean()   # (reference, synthesized)
        loss_extra = ex_loss_weight * lp
        logs["loss_ae_dist_extra"] = float(loss_extra)
        logs["loss_ae_extra"] = float(loss_extra)
        loss = loss_dist + loss_extra
        logs["s_between_mix"] = s_mix.detach()
    else:
        with torch.no_grad():
            z_mix = 0.5 * z[:B] + 0.5 * z[B:]
            z_ref = encode(work, args, slice_between, train=False)
            logs["loss_latent_1"] = float(F.mse_loss(z_mix, z_ref))
        loss = loss_dist
    logs["loss_ae"] = float(loss)
    grads = torch.autograd.grad(loss, [params[k] for k in keys])
    grads = dict(zip(keys, grads))
    for k, v in work.items():                                        # carry BN buffer updates back
        if k not in params:
            state[k] = v
    if adam is not None:
        with torch.no_grad():
            adam.step(state, grads)
    logs["reconstruction"] = out.detach()
    logs["z"] = z.detach()
    if return_grads:
        logs["grads"] = grads
    return logs


# ----------------------------------------------------------------------------------------------
# SSIM / PSNR  (PARITY UNPINNED: scikit-image is an absent, unpinned third-party dependency)
# ----------------------------------------------------------------------------------------------
def _uniform_filter_reflect(a: np.ndarray, win: int) -> np.ndarray:
    """scipy.ndimage.uniform_filter(a, size=win, mode='reflect') restated for 2-D float64 with odd ``win``
    ('reflect' = half-sample symmetric, d c b a | a b c d | d c b a = numpy 'symmetric')."""
    r = win // 2
    p = np.pad(a, r, mode="symmetric")
    c = np.cumsum(np.pad(p, ((1, 0), (0, 0))), axis=0)
    p = (c[win:] - c[:-win]) / win
    c = np.cumsum(np.pad(p, ((0, 0), (1, 0))), axis=1)
    return (c[:, win:] - c[:, :-win]) / win


def ssim_slice(im1: np.ndarray, im2: np.ndarray, win_size: int = 7, data_range: float = 2.0) -> float:
    """skimage.metrics.structural_similarity as the reference invokes it (evaluate/metrics.py:139):
    float images, no data_range argument => legacy dtype range of float = 2.0, win_size 7, uniform filter,
    sample covariance, K1=.01, K2=.03, border of (win-1)//2 cropped before the mean."""
    x, y = im1.astype(np.float64), im2.astype(np.float64)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    ux, uy = _uniform_filter_reflect(x, win_size), _uniform_filter_reflect(y, win_size)
    uxx, uyy, uxy = (_uniform_filter_reflect(x * x, win_size), _uniform_filter_reflect(y * y, win_size),
                     _uniform_filter_reflect(x * y, win_size))
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64))


def psnr_slice(im_true: np.ndarray, im_test: np.ndarray) -> float:
    """skimage.metrics.peak_signal_noise_ratio without data_range (evaluate/metrics.py:188): for float input
    data_range = 1 when im_true.min() >= 0 else 2; 10 log10(R^2 / MSE)."""
    t, s = im_true.astype(np.float64), im_test.astype(np.float64)
    R = 1.0 if t.min() >= 0 else 2.0
    mse = np.mean((t - s) ** 2, dtype=np.float64)
    return float(10 * np.log10(R * R / mse))


def compute_ssim_for_batch(images: np.ndarray, recons: np.ndarray, downsample_steps: Optional[int] = None,
                           data_range: float = 2.0) -> float:
    """evaluate/metrics.py:111-156 (eval_axis=0, normalize=False): mean SSIM over non-original slices."""
    images, recons = images.astype(np.float32), recons.astype(np.float32)
    skip = set(determine_original_sliceids(images.shape[0], downsample_steps).tolist()) if downsample_steps else set()
    res = [ssim_slice(images[z], recons[z], data_range=data_range) for z in range(images.shape[0]) if z not in skip]
    return float(np.mean(np.array(res)))


def compute_psnr_for_batch(images: np.ndarray, recons: np.ndarray, downsample_steps: Optional[int] = None) -> float:
    """evaluate/metrics.py:159-194 (eval_axis=0): mean PSNR over non-original slices, nan/inf dropped."""
    images, recons = images.astype(np.float32), recons.astype(np.float32)
    skip = set(determine_original_sliceids(images.shape[0], downsample_steps).tolist()) if downsample_steps else set()
    res = []
    for z in range(images.shape[0]):
        if z in skip:
            continue
        with np.errstate(divide="ignore"):
            p = psnr_slice(images[z], recons[z])
        if not np.isnan(p) and not np.isinf(p):
            res.append(p)
    return float(np.mean(np.array(res)))


# ----------------------------------------------------------------------------------------------
# VIF (pixel-domain, multi-scale) as the reference computes it: evaluate/vifvec.py:7-63 called with UINT8 images
# (evaluate/metrics.py:72-73,99).  The arithmetic below the reference is scipy.ndimage.gaussian_filter (scipy is
# present here, so the restatement is pinned bit for bit by oracle/make_golden.py::gold_vif against the reference's
# own function).  Quirks that shape the numbers and are reproduced on purpose:
#   * gaussian_filter on a uint8 array returns uint8, and so does every 1-D pass in between: each pass accumulates in
#     float64 (centre tap first, then symmetric pairs from the farthest tap inwards, ni_filters.c NI_Correlate1D) and is
#     truncated to uint8;
#   * `ref * ref`, `mu1 * mu1` and the variance subtractions are uint8 arithmetic and wrap modulo 256.
# ----------------------------------------------------------------------------------------------
VIF_SIGMA_NSQ = 2.0
VIF_EPS = 1e-10


def gaussian_kernel1d(sd: float):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(4 * sd + 0.5) (truncate = 4)."""
    lw = int(4.0 * float(sd) + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum(), lw


def _reflect_index(i: np.ndarray, n: int) -> np.ndarray:
    """scipy 'reflect' (half-sample symmetric: d c b a | a b c d | d c b a), any overhang."""
    i = np.mod(i, 2 * n)
    return np.where(i < n, i, 2 * n - 1 - i)


def gaussian_filter1d_u8(a: np.ndarray, sd: float, axis: int) -> np.ndarray:
    """One pass of scipy.ndimage.gaussian_filter on a uint8 array: float64 accumulation in scipy's order, C cast to uint8."""
    w, lw = gaussian_kernel1d(sd)
    a = np.moveaxis(a, axis, -1)
    n = a.shape[-1]
    idx = np.arange(n)
    x = a.astype(np.float64)
    tmp = x[..., idx] * w[lw]
    for jj in range(-lw, 0):
        left, right = x[..., _reflect_index(idx + jj, n)], x[..., _reflect_index(idx - jj, n)]
        tmp = tmp + (left + right) * w[jj + lw]
    return np.moveaxis(tmp.astype(np.uint8), -1, axis)       # values are in [0, 255]: the cast truncates


def simulate_thick_slices(img3d: np.ndarray, slice_thickness: float) -> np.ndarray:
    """datasets/common_brains.py:37-44: per (y, x) column scipy.ndimage.gaussian_filter1d along z with
    sigma = slice_thickness / 2.355 (FWHM of the slice profile), 'reflect' borders, truncate 4: float64 accumulation in
    scipy's order (centre tap, then symmetric pairs from the farthest tap inwards), result cast to the input dtype."""
    sd = slice_thickness / 2.355
    w, lw = gaussian_kernel1d(sd)
    n = img3d.shape[0]
    idx = np.arange(n)
    x = img3d.astype(np.float64)
    tmp = x * w[lw]
    for jj in range(-lw, 0):
        tmp = tmp + (x[_reflect_index(idx + jj, n)] + x[_reflect_index(idx - jj, n)]) * w[jj + lw]
    return tmp.astype(img3d.dtype)


def gaussian_filter_u8(a: np.ndarray, sd: float) -> np.ndarray:
    """scipy.ndimage.gaussian_filter(a, sd) for a 2-D uint8 array: axis 0 then axis 1, uint8 in between."""
    return gaussian_filter1d_u8(gaussian_filter1d_u8(a, sd, 0), sd, 1)


def vifp_mscale_u8(ref: np.ndarray, dist: np.ndarray, sigma_nsq: float = VIF_SIGMA_NSQ) -> float:
    """evaluate/vifvec.py:7-63 for 2-D uint8 inputs (do_rescale=False)."""
    assert ref.dtype == np.uint8 and dist.dtype == np.uint8
    eps = VIF_EPS
    num, den = 0.0, 0.0
    for scale in range(1, 5):
        N = 2 ** (4 - scale + 1) + 1
        sd = N / 5.0
        if scale > 1:
            ref = gaussian_filter_u8(ref, sd)[::2, ::2]
            dist = gaussian_filter_u8(dist, sd)[::2, ::2]
        mu1, mu2 = gaussian_filter_u8(ref, sd), gaussian_filter_u8(dist, sd)
        mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2                      # uint8, modulo 256
        sigma1_sq = gaussian_filter_u8(ref * ref, sd) - mu1_sq                        # uint8, modulo 256
        sigma2_sq = gaussian_filter_u8(dist * dist, sd) - mu2_sq
        sigma12 = gaussian_filter_u8(ref * dist, sd) - mu1_mu2
        g = sigma12 / (sigma1_sq + eps)
        sv_sq = sigma2_sq - g * sigma12
        m1 = sigma1_sq < eps
        g[m1] = 0
        sv_sq[m1] = sigma2_sq[m1]
        m2 = sigma2_sq < eps
        g[m2] = 0
        sv_sq[m2] = 0
        sv_sq[g < 0] = sigma2_sq[g < 0]
        g[g < 0] = 0
        sv_sq[sv_sq <= eps] = eps
        s1 = sigma1_sq.astype(np.float64)
        num += np.sum(np.log10(1 + g * g * s1 / (sv_sq + sigma_nsq)))
        den += np.sum(np.log10(1 + s1 / sigma_nsq))
    return float(num / den) if den != 0 else float("nan")


def quantize_u8(x: np.ndarray) -> np.ndarray:
    """evaluate/metrics.py:72-73: np.uint8(np.clip(x * 255., 0, 255)) on float32 images (product in float32)."""
    return np.uint8(np.clip(x.astype(np.float32) * 255., 0, 255))


def compute_vif_for_batch(images: np.ndarray, recons: np.ndarray, downsample_steps: Optional[int] = None) -> float:
    """evaluate/metrics.py:65-108 (eval_axis=0, normalize=False): mean VIF over non-original slices, nan/inf dropped."""
    a, b = quantize_u8(images), quantize_u8(recons)
    skip = set(determine_original_sliceids(a.shape[0], downsample_steps).tolist()) if downsample_steps else set()
    res = []
    for z in range(a.shape[0]):
        if z in skip:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            v = vifp_mscale_u8(a[z], b[z])
        if not np.isnan(v) and not np.isinf(v):
            res.append(v)
    return float(np.mean(np.array(res)))


def determine_last_slice(orig_num_slices: int, downsample_steps: int) -> int:
    """evaluate/common.py:36-39."""
    return orig_num_slices - 1 - ((orig_num_slices - 1) % downsample_steps)


def compute_metrics(images_ref: np.ndarray, new_images: np.ndarray, downsample_steps: int) -> Dict[str, float]:
    """evaluate/create_HR_images.py:121-178 for one volume (eval_axis=0, normalize=False, no LPIPS): metrics over all
    slices up to the last synthesised pair, over the synthesised slices only and over the reconstructed ones only."""
    last = determine_last_slice(images_ref.shape[0], downsample_steps) + 1
    r_mask, s_mask = synth_slice_mask(images_ref.shape[0], downsample_steps)
    out = {}
    for tag, sel in (("", slice(None)), ("_synth", s_mask), ("_recon", r_mask)):
        a, b = images_ref[:last][sel], new_images[:last][sel]
        out["ssim" + tag] = compute_ssim_for_batch(a, b)
        out["psnr" + tag] = compute_psnr_for_batch(a, b)
        out["vif" + tag] = compute_vif_for_batch(a, b)
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench / golden generation
# ----------------------------------------------------------------------------------------------
def synthetic_volume(num_slices: int, size: int, seed: int = 1) -> torch.Tensor:
    """uniform [0,1) volume [Z,1,size,size]; CPU generator => identical on every machine."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(num_slices, 1, size, size, generator=g)


def smooth_phantom(num_slices: int, size: int, seed: int = 2, sigma: float = 6.0) -> torch.Tensor:
    """Smooth phantom: separable-Gaussian-blurred noise, slices a running mix .8 prev + .2 new, rescaled to [0,1]."""
    g = torch.Generator().manual_seed(seed)
    r = int(3 * sigma)
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, dtype=torch.float32) / sigma) ** 2)
    k = k / k.sum()
    prev, slices = None, []
    for _ in range(num_slices):
        n = torch.rand(1, 1, size, size, generator=g)
        n = F.conv2d(F.pad(n, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
        n = F.conv2d(F.pad(n, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
        prev = n if prev is None else 0.8 * prev + 0.2 * n
        slices.append(prev)
    v = torch.cat(slices, dim=0)
    return (v - v.min()) / (v.max() - v.min())


def mri_phantom(num_slices: int, size: int, seed: int = 2, sigma: float = 5.0, noise: float = 0.02) -> torch.Tensor:
    """MRI-like phantom [Z,1,size,size]: soft-thresholded smooth noise (bright blobs with edges on a black background,
    adjacent slices correlated) + a little pixel noise -- the training / test images of the reference-trained checkpoint
    (oracle/make_golden.py::gold_trained)."""
    v = smooth_phantom(num_slices, size, seed=seed, sigma=sigma)
    g = torch.Generator().manual_seed(seed + 7919)
    fine = smooth_phantom(num_slices, size, seed=seed + 1, sigma=1.5)
    img = ((v - 0.38) * 2.6).clamp(0, 1) * (0.75 + 0.25 * fine)
    return (img + noise * torch.rand(img.shape, generator=g)).clamp(0, 1)


def default_args(width=128, latent_width=32, latent=128, depth=32) -> dict:
    return {"width": width, "latent_width": latent_width, "latent": latent, "depth": depth, "colors": 1,
            "n_res_block": None, "use_batchnorm": True, "use_sigmoid": True, "device": "cpu"}


def calibrated_state(args: dict, seed: int = 892372, calib_seed: int = 5) -> "OrderedDict[str, torch.Tensor]":
    """A 'trained-like' synthetic checkpoint for parity tests.  The literal reference init gives eval-mode
    latents of order 1e-6 (tiny init std, identity BN), which would make every parity check vacuous.  Here the
    conv weights of ``init_state`` are rescaled to He gain, biases / BN beta get small seeded values, and BN
    running stats are set to the batch statistics of a seeded phantom (momentum 1), so eval-mode activations
    are O(1) in every layer.  Deterministic on CPU; the reference model loads it via ``load_state_dict``."""
    st = init_state(args, seed=seed)
    g = torch.Generator().manual_seed(calib_seed)
    for k in list(st.keys()):
        v = st[k]
        if k.endswith(".weight") and v.dim() == 4:
            cout, cin, kh, kw = v.shape
            st[k] = v * float(np.sqrt(2.0 / (cin * kh * kw)) * np.sqrt(1.04 * cout * cin * kh))
        elif k.endswith(".weight") and v.dim() == 1:
            st[k] = torch.where(v.abs() < 0.3, torch.full_like(v, 0.3) * torch.sign(v + 1e-12), v)
        elif k.endswith(".bias"):
            st[k] = (torch.rand(v.shape, generator=g) - 0.5) * 0.2
    x = 0.7 * smooth_phantom(6, args["width"], seed=calib_seed) + 0.3 * synthetic_volume(6, args["width"], seed=calib_seed + 1)
    scales = num_scales(args["width"], args["latent_width"])
    with torch.no_grad():
        z = _run(encoder_spec(scales, args["depth"], args["latent"]), "enc", st, x, True, None, momentum=1.0)
        _run(decoder_spec(scales, args["depth"], args["latent"]), "dec", st, z, True, None, momentum=1.0)
    return st


def randomize_bn_stats(state, seed: int = 7) -> None:
    """Give BN running stats / affine non-trivial values so eval-mode BN is not the identity (deterministic)."""
    g = torch.Generator().manual_seed(seed)
    for k in list(state.keys()):
        if k.endswith("running_mean"):
            state[k] = (torch.rand(state[k].shape, generator=g) - 0.5) * 0.2
        elif k.endswith("running_var"):
            state[k] = 0.5 + torch.rand(state[k].shape, generator=g)
