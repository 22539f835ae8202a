"""LPIPS-VGG v0.1 ('net-lin', 'vgg') on the sm_100a kernels -- drop-in for ``lpips.perceptual.PerceptualLoss``
(lpips/perceptual.py:6-33 -> lpips/dist_model.py:100-108 -> lpips/networks_basic.py:63-91).

The VGG16 trunk (lpips/pretrained_networks.py:97-135) runs on the tcgen05 conv kernel (ReLU and the 2x2 max-pool fused
into the epilogue, full-resolution tap + pooled tensor written by the same launch); conv1_1 carries the
``2x-1`` / ScalingLayer / 1->3-channel broadcast; the distance head (channel-L2 normalise, squared difference, 1x1
``lin`` conv, spatial mean, sum over the five taps) is one fused reduction kernel per tap, forward and backward.
Only the trunk's data gradient exists (parameters are frozen, pretrained_networks.py:117-119), and only for the
synthesized branch.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops, ops_train as T

VGG_CFG = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512),
           (512, 512), (512, 512), (512, 512), (512, 512)]
POOL_AFTER = (1, 3, 6, 9)                 # conv indices followed by MaxPool2d(2)
TAPS = (1, 3, 6, 9, 12)                   # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
SHIFT = (-.030, -.088, -.188)             # ScalingLayer buffers, lpips/networks_basic.py:96-97
SCALE = (.458, .448, .450)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lpips_vgg_lin_v0_1.npz")


def vgg16_random_init(seed: Optional[int] = None) -> List[torch.Tensor]:
    """The 13 conv weight/bias pairs as ``torchvision.models.vgg16(weights=None)`` initialises them -- TEST / BENCH ONLY
    (explicit opt-in through ``random_init_seed``): a randomly initialised trunk is not LPIPS.  Same RNG stream as
    torchvision, so a seed reproduces the oracle's ``init_vgg``."""
    if seed is not None:
        torch.manual_seed(seed)
    convs = [nn.Conv2d(cin, cout, 3, padding=1) for cin, cout in VGG_CFG]
    for fin, fout in ((512 * 7 * 7, 4096), (4096, 4096), (4096, 1000)):
        nn.Linear(fin, fout)                      # torchvision builds the classifier before its init loop
    out = []
    for c in convs:
        nn.init.kaiming_normal_(c.weight, mode="fan_out", nonlinearity="relu")
        nn.init.constant_(c.bias, 0)
        out += [c.weight.data, c.bias.data]
    return out


VGG16_HUB_FILES = ("vgg16-397923af.pth",)        # torchvision IMAGENET1K_V1 = what vgg16(pretrained=True) downloads


def load_vgg16_features(path: str) -> List[torch.Tensor]:
    """The 13 conv weight/bias pairs of a torchvision VGG16 checkpoint (``features.N.weight`` keys, or the bare
    ``features`` Sequential, or the reference's ``slice{k}.N`` naming, lpips/pretrained_networks.py:104-116)."""
    sd = torch.load(os.path.expanduser(path), map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    ws = [(k, v) for k, v in sd.items() if k.endswith(".weight") and torch.is_tensor(v) and v.dim() == 4]
    if len(ws) < 13:
        raise RuntimeError("aesr_b200 LPIPS: %s holds %d conv filters, a VGG16 trunk has 13" % (path, len(ws)))
    out = []
    for (k, w), (cin, cout) in zip(ws[:13], VGG_CFG):
        if tuple(w.shape) != (cout, cin, 3, 3):
            raise RuntimeError("aesr_b200 LPIPS: %s in %s has shape %s, expected %s" % (k, path, tuple(w.shape),
                                                                                        (cout, cin, 3, 3)))
        out += [w.float(), sd[k[:-len("weight")] + "bias"].float()]
    return out


def resolve_vgg16_state(vgg_state=None, vgg_weights: Optional[str] = None, random_init_seed: Optional[int] = None
                        ) -> List[torch.Tensor]:
    """Where the LPIPS trunk's filters come from, in this order: explicit tensors; a checkpoint path (argument, then
    ``AESR_VGG16_WEIGHTS``); torchvision's ImageNet file in the torch hub cache (what the reference's
    ``vgg16(pretrained=True)`` leaves there, lpips/pretrained_networks.py:100); a SEEDED random trunk only on explicit
    request (``random_init_seed`` / ``AESR_LPIPS_RANDOM_INIT_SEED`` -- identical on every rank).  Otherwise raise: a loss
    computed on an unseeded random trunk differs per run and per data-parallel rank and is not LPIPS."""
    if vgg_state is not None:
        return list(vgg_state)
    path = vgg_weights or os.environ.get("AESR_VGG16_WEIGHTS")
    if path:
        return load_vgg16_features(path)
    hub = os.path.join(torch.hub.get_dir(), "checkpoints")
    for name in VGG16_HUB_FILES:
        if os.path.isfile(os.path.join(hub, name)):
            return load_vgg16_features(os.path.join(hub, name))
    if random_init_seed is None and os.environ.get("AESR_LPIPS_RANDOM_INIT_SEED"):
        random_init_seed = int(os.environ["AESR_LPIPS_RANDOM_INIT_SEED"])
    if random_init_seed is not None:
        state = torch.random.get_rng_state()
        try:
            return vgg16_random_init(int(random_init_seed))
        finally:
            torch.random.set_rng_state(state)
    raise RuntimeError(
        "aesr_b200 LPIPS: no VGG16 weights.  The reference uses torchvision's ImageNet-pretrained vgg16 "
        "(lpips/pretrained_networks.py:100); give its checkpoint as args['vgg_weights'] / PerceptualLoss(vgg_weights=...) / "
        "AESR_VGG16_WEIGHTS, or place %s under %s.  For tests and benchmarks only, a seeded random trunk can be requested "
        "with args['lpips_random_init_seed'] / AESR_LPIPS_RANDOM_INIT_SEED." % (VGG16_HUB_FILES[0], hub))


class PerceptualLoss(nn.Module):
    """Same constructor / forward signature as the reference class; ``model='net-lin', net='vgg'`` only."""

    def __init__(self, model="net-lin", net="vgg", colorspace="rgb", spatial=False, use_gpu=True, gpu_ids=[0],
                 vgg_state: Optional[List[torch.Tensor]] = None, device=None, act_dtype=None,
                 vgg_weights: Optional[str] = None, random_init_seed: Optional[int] = None):
        super().__init__()
        if model != "net-lin" or net not in ("vgg", "vgg16") or spatial:
            raise NotImplementedError("aesr_b200 LPIPS: only model='net-lin', net='vgg', spatial=False are on the hot "
                                      "path (kwatsch/base_trainer.py:41-43)")
        dev = torch.device(device if device is not None else "cuda:%d" % int(gpu_ids[0]))
        state = resolve_vgg16_state(vgg_state, vgg_weights, random_init_seed)
        self.weights = nn.ParameterList([nn.Parameter(t.detach().clone().float().to(dev), requires_grad=False)
                                         for t in state])
        lins = np.load(_DATA)
        self.lins = nn.ParameterList([nn.Parameter(torch.from_numpy(lins["lin%d" % k]).reshape(-1).to(dev),
                                                   requires_grad=False) for k in range(5)])
        self.dtype = act_dtype or ops.DEFAULT_DTYPE
        self._fwd_packed = None
        self._bwd_packed = None

    # ---------------------------------------------------------------- packed filters (frozen => packed once)
    def _packs(self):
        if self._fwd_packed is None:
            self._fwd_packed = [None] + [ops.pack_conv3x3_weight(self.weights[2 * i], dtype=self.dtype)
                                         for i in range(1, 13)]
            self._bwd_packed = [None] + [ops.pack_conv3x3_weight(self.weights[2 * i], transpose_flip=True,
                                                                 dtype=T.GRAD_DTYPE) for i in range(1, 13)]
        return self._fwd_packed, self._bwd_packed

    def _trunk(self, img: torch.Tensor, normalize: bool):
        """img fp32 [N,1,H,W] -> list of the 13 post-ReLU activations (NHWC 16-bit, full resolution before pools)."""
        fwd, _ = self._packs()
        h = T.vgg_conv1_fwd(img, self.weights[0], self.weights[1], SHIFT, SCALE, normalize, self.dtype)
        acts = [h]
        for i in range(1, 13):
            if i in POOL_AFTER:
                full, h = ops.conv3x3(h, fwd[i], self.weights[2 * i + 1], act=ops.ACT_RELU,
                                      out_mode=ops.OUT_SAME_MAXPOOL2)
                acts.append(full)
            else:
                h = ops.conv3x3(h, fwd[i], self.weights[2 * i + 1], act=ops.ACT_RELU)
                acts.append(h)
        return acts

    @torch.no_grad()
    def forward(self, pred: torch.Tensor, target: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[N,1,H,W] x 2 -> [N,1,1,1] distances (no autograd graph; the training engine uses value_and_grad)."""
        n = pred.shape[0]
        imgs = torch.cat([target.detach().float(), pred.detach().float()], dim=0).contiguous()
        acts = self._trunk(imgs, normalize)
        val = torch.zeros(n, dtype=torch.float32, device=imgs.device)
        for k, ci in enumerate(TAPS):
            T.lpips_head(acts[ci][:n], acts[ci][n:], self.lins[k], val)
        return val.view(n, 1, 1, 1)

    @torch.no_grad()
    def reference_features(self, reference: torch.Tensor, normalize: bool = True):
        """Trunk activations of the reference images alone: they depend on the batch only, not on the autoencoder, so the
        training engine computes them on a second stream while the autoencoder's forward pass runs."""
        return self._trunk(reference.detach().float().contiguous(), normalize)

    @torch.no_grad()
    def value_and_grad(self, reference: torch.Tensor, synthesized: torch.Tensor, upstream: torch.Tensor,
                       normalize: bool = True, grad_out: Optional[torch.Tensor] = None, ref_acts=None):
        """Per-image distances [N] and d(sum_n upstream[n] * val[n]) / d synthesized  (fp32 [N,1,H,W]).
        ``ref_acts``: ``reference_features(reference)`` computed earlier (then only the synthesized half runs here)."""
        n = reference.shape[0]
        _, bwd = self._packs()
        if ref_acts is None:
            imgs = torch.cat([reference.detach().float(), synthesized.detach().float()], dim=0).contiguous()
            acts = self._trunk(imgs, normalize)
            ref = [a[:n] for a in acts]
            syn = [a[n:] for a in acts]
            dev = imgs.device
        else:
            ref = ref_acts
            syn = self._trunk(synthesized.detach().float().contiguous(), normalize)
            dev = synthesized.device
        val = torch.zeros(n, dtype=torch.float32, device=dev)
        g_tap = {}
        for k, ci in enumerate(TAPS):
            g_tap[ci] = T.lpips_head(ref[ci], syn[ci], self.lins[k], val, upstream, want_grad=True)
        # walk the trunk backwards on the synthesized half
        g = T.maxpool_bwd(syn[12], None, g_tap[12])                       # relu'(c5_3) * tap gradient
        for i in range(12, 0, -1):
            if (i - 1) in POOL_AFTER:                                     # conv i reads the pooled output of conv i-1
                d_pooled = ops.conv3x3(g, bwd[i], None)
                g = T.maxpool_bwd(syn[i - 1], d_pooled, g_tap.get(i - 1))
            else:
                g = ops.conv3x3(g, bwd[i], None, mul_src=syn[i - 1], mul_mode=ops.MUL_RELU_GRAD)
        dimg = T.vgg_conv1_bwd(g, self.weights[0], SCALE, normalize, out=grad_out)
        return val, dimg
