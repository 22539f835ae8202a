"""``VanillaACAI`` -- drop-in for the reference model class (networks/acai_vanilla.py:112-138).

Same constructor (one ``args`` dict), same ``encode / decode / forward`` signatures, same ``state_dict`` keys and
shapes (``enc.N.weight`` ... SURVEY.md App. D), same random initialisation stream (``Initializer``,
networks/acai_vanilla.py:39-46) -- but no torch.nn compute modules: the ``enc`` / ``dec`` containers only *hold*
parameters; all arithmetic runs in the sm_100a kernels of libaesr_b200.so (``ops.py``).  There is no CPU path.

Layer structure follows ``Encoder`` (:49-72) and ``Decoder`` (:75-102) with use_batchnorm / use_upsample /
use_sigmoid as the ae_combined configs set them (networks/net_config.py:28-29).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import ops

LEAKY_SLOPE = 0.01
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class _Holder(nn.Module):
    def forward(self, *a, **k):          # pragma: no cover
        raise RuntimeError("parameter holder only: computation runs in the aesr_b200 CUDA kernels via "
                           "VanillaACAI.encode/decode")


class ConvHolder(_Holder):
    """Parameters of an nn.Conv2d(cin, cout, k, padding=p); reset_parameters draws like torch's so that the RNG
    stream (and therefore the reference Initializer's result) is reproduced bit for bit."""

    def __init__(self, cin: int, cout: int, k: int, padding: int):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.padding = cin, cout, k, padding
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(cin * k * k)
        nn.init.uniform_(self.bias, -bound, bound)


class BatchNormHolder(_Holder):
    def __init__(self, c: int):
        super().__init__()
        self.num_features, self.eps, self.momentum = c, BN_EPS, BN_MOMENTUM
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Marker(_Holder):
    """Parameter-free position in the reference nn.Sequential (LeakyReLU / AvgPool2d / Upsample / Sigmoid)."""

    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind

    def extra_repr(self):
        return self.kind


def _initializer(layers, slope: float = 0.2):
    """networks/acai_vanilla.py:39-46: every layer with a weight (convs AND BatchNorm) ~ N(0, std),
    std = 1/sqrt((1+slope^2) * prod(shape[:-1])); biases zero."""
    for layer in layers:
        if hasattr(layer, "weight"):
            w = layer.weight.data
            std = 1 / np.sqrt((1 + slope ** 2) * np.prod(w.shape[:-1]))
            w.normal_(std=std)
        if hasattr(layer, "bias"):
            layer.bias.data.zero_()


def build_encoder(scales, depth, latent, colors) -> nn.Sequential:
    layers: List[nn.Module] = [ConvHolder(colors, depth, 1, 1)]
    kp = depth
    for s in range(scales):
        k = depth << s
        layers += [ConvHolder(kp, k, 3, 1), Marker("leaky"), ConvHolder(k, k, 3, 1), Marker("leaky"),
                   BatchNormHolder(k), Marker("avgpool2")]
        kp = k
    k = depth << scales
    layers += [ConvHolder(kp, k, 3, 1), Marker("leaky"), ConvHolder(k, latent, 3, 1)]
    _initializer(layers)
    return nn.Sequential(*layers)


def build_decoder(scales, depth, latent, colors) -> nn.Sequential:
    layers: List[nn.Module] = []
    kp = latent
    for s in range(scales - 1, -1, -1):
        k = depth << s
        layers += [ConvHolder(kp, k, 3, 1), Marker("leaky"), ConvHolder(k, k, 3, 1), Marker("leaky"),
                   BatchNormHolder(k), Marker("upsample2")]
        kp = k
    layers += [ConvHolder(kp, depth, 3, 1), Marker("leaky"), ConvHolder(depth, colors, 3, 1), Marker("sigmoid")]
    _initializer(layers)
    return nn.Sequential(*layers)


class _PackedCache:
    """Derived device tensors (bf16 packed filters, folded eval-BN affine) keyed on the parameters' versions."""

    def __init__(self):
        self._store = {}

    def clear(self):
        self._store = {}

    def get(self, key, tensors, build):
        ver = tuple((t.data_ptr(), t._version) for t in tensors)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = build()
        self._store[key] = (ver, val)
        return val


class VanillaACAI(nn.Module):
    def __init__(self, args: dict):
        super().__init__()
        scales = int(round(math.log(args["width"] // args["latent_width"], 2)))
        args.setdefault("n_res_block", None)
        args.setdefault("use_batchnorm", False)
        args.setdefault("use_sigmoid", False)
        args.setdefault("gpu_ids", [0])
        if args["n_res_block"] is not None:
            raise NotImplementedError("aesr_b200: n_res_block is not part of the ae_combined hot path (None in "
                                      "every NetworkConfig, networks/net_config.py:26)")
        if not (args["use_batchnorm"] and args["use_sigmoid"]):
            raise NotImplementedError("aesr_b200 implements the ae_combined configuration: use_batchnorm=True, "
                                      "use_sigmoid=True (networks/net_config.py:28-29)")
        if args.get("colors", 1) != 1:
            raise NotImplementedError("aesr_b200: colors must be 1 (single-channel MRI slices)")
        if args["depth"] != 32:
            raise NotImplementedError("aesr_b200: depth must be 32 (channel counts are tiled for the tensor cores)")
        self.scales, self.depth, self.latent = scales, args["depth"], args["latent"]
        self.enc = build_encoder(scales, args["depth"], args["latent"], 1).to(args["device"])
        self.dec = build_decoder(scales, args["depth"], args["latent"], 1).to(args["device"])
        self._cache = _PackedCache()
        # Inference pipelines use the algebraic folds (stem = enc.0 o enc.1, Upsample folded into the next conv, decoder
        # head fused into dec.12's epilogue).  False = one kernel per reference layer (A/B measurements, tests).
        self.fused_inference = os.environ.get("AESR_FUSED", "1") != "0"
        # Volume synthesis: interpolate BEHIND dec.0 (dec.0 is linear; it then runs once per low-res slice).
        self.linear_fold = os.environ.get("AESR_LINFOLD", "1") != "0"

    # ------------------------------------------------------------------ public API (reference signatures)
    def forward(self, img):
        return self.decode(self.encode(img))

    def _train_forward(self):
        """Forward-only train-mode helper (batch-statistics BN, running stats updated).  Gradients are never built by
        autograd here: optimisation goes through ``training.engine.TrainEngine.step`` (the trainers' ``train()``)."""
        if getattr(self, "_tf", None) is None:
            from ..training.engine import TrainForward
            self._tf = TrainForward(self)
        self._tf._pack_all()
        return self._tf

    def invalidate_cache(self):
        """Parameters / BN buffers were updated in place by our kernels (no torch version bump): drop derived tensors."""
        self._cache.clear()

    @staticmethod
    def _no_autograd(t, what):
        """The reference module is differentiable through autograd (networks/acai_vanilla.py:130-138); this one runs
        hand-written kernels and builds no graph.  Gradients exist through ``training.engine.TrainEngine.step`` (the
        trainers' ``train()``).  A caller that expects autograd gradients must hear about it instead of silently
        getting a detached tensor."""
        if torch.is_grad_enabled() and torch.is_tensor(t) and t.requires_grad:
            raise RuntimeError("aesr_b200: VanillaACAI.%s received a tensor that requires grad with autograd enabled; this "
                               "module builds no autograd graph (its backward pass is TrainEngine.step / trainer.train()). "
                               "Call it under torch.no_grad() or detach the input." % what)

    def encode(self, img):
        self._no_autograd(img, "encode")
        with torch.no_grad():
            return self._encode(img)

    def decode(self, z):
        self._no_autograd(z, "decode")
        with torch.no_grad():
            return self._decode(z)

    def _encode(self, img):
        if self.training:
            z, _, _ = self._train_forward().encode_train(img.detach().float().contiguous(), save=False)
            self.invalidate_cache()
            return z
        return self.encode_eval(img)

    def _decode(self, z):
        if self.training:
            z = z.detach().float().contiguous()
            m = z.shape[0]
            idx = torch.arange(m, dtype=torch.int32, device=z.device)
            one = torch.ones(m, dtype=torch.float32, device=z.device)
            z16 = ops.lerp_latents(z, idx, torch.full_like(idx, -1), one, one)
            out, _ = self._train_forward().decode_train(z16, save=False)
            self.invalidate_cache()
            return out
        return self.decode_eval(z)

    # ------------------------------------------------------------------ eval-mode (inference) pipelines
    def _packed(self, conv: ConvHolder) -> torch.Tensor:
        return self._cache.get(("w", id(conv)), [conv.weight], lambda: ops.pack_conv3x3_weight(conv.weight))

    def _bn_affine(self, bn: BatchNormHolder):
        def build():
            with torch.no_grad():       # tiny per-channel fold (C <= 256 values), host-side plumbing
                scale = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
                shift = bn.bias.double() - bn.running_mean.double() * scale
                return scale.float().contiguous(), shift.float().contiguous()
        return self._cache.get(("bn", id(bn)), [bn.weight, bn.bias, bn.running_mean, bn.running_var], build)

    def _packed_up2(self, conv: ConvHolder) -> torch.Tensor:
        return self._cache.get(("wup", id(conv)), [conv.weight], lambda: ops.pack_conv3x3_weight_up2fold(conv.weight))

    def _stem(self):
        c0, c1 = self.enc[0], self.enc[1]
        return self._cache.get(("stem",), [c0.weight, c0.bias, c1.weight, c1.bias],
                               lambda: ops.stem_host_params(*ops.stem_fold(c0.weight, c0.bias, c1.weight), c1.bias))

    def _head_w_host(self, conv: ConvHolder):
        """fp32 [9,32] head filter in host memory (one device->host copy per parameter version)."""
        return self._cache.get(("head_host", id(conv)), [conv.weight],
                               lambda: conv.weight.detach()[0].permute(1, 2, 0).reshape(9, -1).float().cpu().contiguous())

    def _head_w16(self, conv: ConvHolder):
        """16-bit [16,32] head filter on the device (tensor-core head inside the fused decoder tail)."""
        return self._cache.get(("head16", id(conv), ops.DEFAULT_DTYPE), [conv.weight],
                               lambda: ops.pack_head_w16(conv.weight.detach()[0].permute(1, 2, 0).reshape(9, -1).float()))

    def _head_w(self, conv: ConvHolder):
        def build():
            with torch.no_grad():
                return (conv.weight.detach()[0].permute(1, 2, 0).reshape(9, -1).contiguous().float(),
                        conv.bias.detach().float().contiguous())
        return self._cache.get(("head", id(conv)), [conv.weight, conv.bias], build)

    @torch.no_grad()
    def encode_eval(self, img: torch.Tensor, want_nhwc: bool = False, nhwc_only: bool = False):
        """[N,1,H,W] fp32 -> z [N,latent,h,w] fp32 (eval-mode BN).  ``want_nhwc`` also returns the 16-bit NHWC copy;
        ``nhwc_only`` returns ONLY the 16-bit NHWC latent (volume synthesis: the fp32 NCHW tensor is never needed)."""
        x = img.detach().float().contiguous()
        enc = self.enc
        fused = self.fused_inference
        if fused:
            a = ops.stem(x, self._stem())
        else:
            a = ops.e0(x, enc[0].weight.detach().reshape(-1).contiguous(), enc[0].bias.detach())
        i = 1
        for s in range(self.scales):
            c1, c2, bn = enc[i], enc[i + 2], enc[i + 4]
            if not (fused and s == 0):
                a = ops.conv3x3(a, self._packed(c1), c1.bias.detach(), act=ops.ACT_LEAKY)
            sc, sh = self._bn_affine(bn)
            a = ops.conv3x3(a, self._packed(c2), c2.bias.detach(), act=ops.ACT_LEAKY, scale=sc, shift=sh,
                            out_mode=ops.OUT_AVGPOOL2)
            i += 6
        c1, c2 = enc[i], enc[i + 2]
        a = ops.conv3x3(a, self._packed(c1), c1.bias.detach(), act=ops.ACT_LEAKY)
        if nhwc_only:
            return ops.conv3x3(a, self._packed(c2), c2.bias.detach(), act=ops.ACT_NONE)
        return ops.conv3x3(a, self._packed(c2), c2.bias.detach(), act=ops.ACT_NONE, out_mode=ops.OUT_NCHW_F32,
                           want_out2=want_nhwc)

    @torch.no_grad()
    def decode_pre_eval(self, z16: torch.Tensor) -> torch.Tensor:
        """dec.0 WITHOUT bias / activation on 16-bit NHWC latents -> un-rounded fp32 NHWC pre-activations.  dec.0 is
        linear, so interpolating these (``ops.lerp_pairs_act``: blend, + bias, LeakyReLU) equals dec.0 + LeakyReLU of
        the interpolated latent (networks/acai_vanilla.py:87-88 after generate_hr_volumes.py:88) in real arithmetic."""
        return ops.conv3x3(z16, self._packed(self.dec[0]), None, act=ops.ACT_NONE, out_mode=ops.OUT_SAME_F32)

    @torch.no_grad()
    def decode_nhwc_eval(self, a: torch.Tensor, out: Optional[torch.Tensor] = None,
                         out_image_stride: Optional[int] = None,
                         out_index: Optional[torch.Tensor] = None, after_first: bool = False) -> torch.Tensor:
        """decoder on an NHWC 16-bit latent batch; images optionally written strided into ``out``.  ``after_first``:
        ``a`` is already the output of dec.0 + LeakyReLU (``decode_pre_eval`` + ``ops.lerp_pairs_act``)."""
        dec = self.dec
        if self.fused_inference:
            return self._decode_nhwc_fused(a, out, out_image_stride, out_index, after_first)
        assert not after_first, "after_first needs the fused inference pipeline"
        i = 0
        for _ in range(self.scales):
            c1, c2, bn = dec[i], dec[i + 2], dec[i + 4]
            a = ops.conv3x3(a, self._packed(c1), c1.bias.detach(), act=ops.ACT_LEAKY)
            sc, sh = self._bn_affine(bn)
            a = ops.conv3x3(a, self._packed(c2), c2.bias.detach(), act=ops.ACT_LEAKY, scale=sc, shift=sh,
                            out_mode=ops.OUT_UP2)
            i += 6
        c1, c2 = dec[i], dec[i + 2]
        a = ops.conv3x3(a, self._packed(c1), c1.bias.detach(), act=ops.ACT_LEAKY)
        w9c, b = self._head_w(c2)
        return ops.head(a, w9c, b, out=out, out_image_stride=out_image_stride, sigmoid=True, out_index=out_index)

    def _decode_nhwc_fused(self, a, out, out_image_stride, out_index, after_first=False):
        """Same decoder with every nn.Upsample folded into the conv that follows it (the producer stores the LOW-res
        tensor, the consumer is a low-res conv with 4 phase blocks + depth-to-space) and dec.12 -> dec.14 -> sigmoid as
        one tensor-core kernel + a 20 B/px gather."""
        dec = self.dec
        i = 0
        for s in range(self.scales):
            c1, c2, bn = dec[i], dec[i + 2], dec[i + 4]
            if s == 0:
                if not after_first:
                    a = ops.conv3x3(a, self._packed(c1), c1.bias.detach(), act=ops.ACT_LEAKY)
            else:       # input is the previous block's low-res output: Upsample folded into this conv
                a = ops.conv3x3(a, self._packed_up2(c1), c1.bias.detach(), act=ops.ACT_LEAKY, out_mode=ops.OUT_SHUFFLE2)
            sc, sh = self._bn_affine(bn)
            a = ops.conv3x3(a, self._packed(c2), c2.bias.detach(), act=ops.ACT_LEAKY, scale=sc, shift=sh)
            i += 6
        c1, c2 = dec[i], dec[i + 2]
        w9c, b = self._head_w(c2)
        part = ops.conv3x3_up2_head(a, self._packed_up2(c1), c1.bias.detach(), self._head_w_host(c2), act=ops.ACT_LEAKY,
                                    head_w16=self._head_w16(c2) if ops.HEAD_ON_TENSOR_CORES else None)
        return ops.head_gather(part, b, out=out, out_image_stride=out_image_stride, sigmoid=True, out_index=out_index)

    @torch.no_grad()
    def decode_eval(self, z: torch.Tensor) -> torch.Tensor:
        z = z.detach().float().contiguous()
        m = z.shape[0]
        idx = torch.arange(m, dtype=torch.int32, device=z.device)
        neg = torch.full((m,), -1, dtype=torch.int32, device=z.device)
        one = torch.ones(m, dtype=torch.float32, device=z.device)
        a = ops.lerp_latents(z, idx, neg, one, one)
        return self.decode_nhwc_eval(a)
