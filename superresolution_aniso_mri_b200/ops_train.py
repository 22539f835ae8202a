"""Tensor-level wrappers of the training-step entry points (see include/aesr_b200.h).  Gradients are bf16 NHWC."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .ops import _dev, _ptr, _stream, _timed, dt_code

BN_SAME, BN_POOL, BN_UP = 0, 1, 2
GRAD_DTYPE = torch.bfloat16
_F3 = ctypes.c_float * 3


def bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps, count1: float = 0.0):
    """-> (scale, shift, mean, invstd); running stats updated in place (torch BatchNorm2d.train() semantics).
    ``count1`` > 0: ``stats`` is [2][2C] (two passes of a merged batch), the outputs are [2, C] and the running statistics
    are updated once per pass in order."""
    lib = _dev(stats)
    c = gamma.numel()
    passes = 2 if count1 > 0 else 1
    assert stats.numel() == passes * 2 * c
    out = torch.empty(4, passes * c, dtype=torch.float32, device=stats.device)
    _lib.check(lib.aesr_bn_finalize(stats.data_ptr(), float(count), float(count1), passes, gamma.data_ptr(), beta.data_ptr(),
                                    _ptr(running_mean), _ptr(running_var), float(momentum), float(eps),
                                    out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), c,
                                    _stream(stats)), "bn_finalize")
    return out[0], out[1], out[2], out[3]


def bn_apply(a, scale, shift, mode, split: int = 0):
    lib = _dev(a)
    n, h, w, c = a.shape
    ho, wo = (h // 2, w // 2) if mode == BN_POOL else (2 * h, 2 * w) if mode == BN_UP else (h, w)
    out = torch.empty((n, ho, wo, c), dtype=a.dtype, device=a.device)
    with _timed("bn_apply"):
        _lib.check(lib.aesr_bn_apply(a.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), n, h, w, c, mode,
                                     dt_code(a.dtype), int(split), _stream(a)), "bn_apply")
    return out


def bn_bwd(dnext, a, mean, invstd, gamma, dgamma, dbeta, mode, slope=0.01, sync_world: int = 1, split: int = 0,
           dbias_conv: Optional[torch.Tensor] = None):
    """dnext bf16 (grad of pooled / upsampled BN output), a saved activation -> g bf16 [N,H,W,C].
    ``split``: images >= split are the second pass of a merged batch (mean / invstd [2, C]).
    ``dbias_conv`` [C] += sum over pixels of g (bias gradient of the producing conv; not in SyncBN mode).
    sync_world > 1 (SyncBN parity mode): the per-channel sums are all-reduced between the reduce and apply kernels."""
    lib = _dev(a)
    n, h, w, c = a.shape
    assert dnext.dtype == GRAD_DTYPE and dnext.is_contiguous()
    passes = 2 if 0 < split < n else 1
    assert mean.numel() >= passes * c
    g = torch.empty((n, h, w, c), dtype=GRAD_DTYPE, device=a.device)
    sums = torch.empty(passes * 2 * c, dtype=torch.float32, device=a.device)
    n0 = split if passes == 2 else n
    assert dbias_conv is None or sync_world <= 1

    def call(phase, count, count1):
        _lib.check(lib.aesr_bn_bwd(dnext.data_ptr(), a.data_ptr(), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                   sums.data_ptr(), float(slope), g.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                   _ptr(dbias_conv), n, h, w, c, mode, dt_code(a.dtype), phase, float(count),
                                   float(count1), int(split), _stream(a)), "bn_bwd")
    with _timed("bn_bwd"):
        if sync_world > 1:
            import torch.distributed as dist
            call(1, 0, 0)
            dist.all_reduce(sums)
            call(2, n0 * h * w * sync_world, (n - n0) * h * w * sync_world)
            # dgamma / dbeta were accumulated from the GLOBAL sums on every rank; the gradient all-reduce averages them
        else:
            call(0, 0, 0)
    return g


def mse(a, b, loss_acc, want_grad=False, grad_scale=1.0, grad_out: Optional[torch.Tensor] = None):
    lib = _dev(a)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    assert a.numel() == b.numel()
    d = grad_out if grad_out is not None else (torch.empty_like(a) if want_grad else None)
    assert d is None or (d.dtype == torch.float32 and d.is_contiguous() and d.numel() == a.numel())
    _lib.check(lib.aesr_mse(a.data_ptr(), b.data_ptr(), a.numel(), loss_acc.data_ptr(), _ptr(d), float(grad_scale),
                            _stream(a)), "mse")
    return d


def head_bwd(dout, out, a_in, w9c, dw9c, dbias, slope=0.01, dbias_in: Optional[torch.Tensor] = None):
    """``dbias_in`` [C] += per-channel sums of the returned gradient (bias gradient of the conv that produced a_in)."""
    lib = _dev(a_in)
    n, h, w, c = a_in.shape
    g = torch.empty((n, h, w, c), dtype=GRAD_DTYPE, device=a_in.device)
    with _timed("head_bwd"):
        _lib.check(lib.aesr_head_bwd(dout.data_ptr(), out.data_ptr(), a_in.data_ptr(), w9c.data_ptr(), g.data_ptr(),
                                     dw9c.data_ptr(), dbias.data_ptr(), _ptr(dbias_in), n, h, w, c, float(slope),
                                     dt_code(a_in.dtype), _stream(a_in)), "head_bwd")
    return g


def e0_bwd(g, x, dw, db):
    lib = _dev(g)
    n, _, h, w = x.shape
    _lib.check(lib.aesr_e0_bwd(g.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), n, h, w, g.shape[-1],
                               _stream(g)), "e0_bwd")


WGRAD_ALGO = 0          # 0 auto (tensor cores when the channel counts allow), 1 tcgen05, 2 CUDA cores


def wgrad3x3(g, x, dW, dbias: Optional[torch.Tensor], algo: Optional[int] = None):
    """dW [Cout,Cin,3,3] fp32 += ..., dbias [Cout] += ...; g bf16 [N,H,W,Cout], x 16-bit [N,H,W,Cin]."""
    lib = _dev(g)
    n, h, w, cout = g.shape
    cin = x.shape[-1]
    assert g.dtype == GRAD_DTYPE and dW.shape == (cout, cin, 3, 3) and dW.is_contiguous()
    with _timed("wgrad3x3", 2.0 * n * h * w * 9 * cin * cout):
        _lib.check(lib.aesr_wgrad3x3(g.data_ptr(), x.data_ptr(), dW.data_ptr(), _ptr(dbias), n, h, w, cin, cout,
                                     dt_code(x.dtype), WGRAD_ALGO if algo is None else int(algo), _stream(g)),
                   "wgrad3x3")


def mix_bwd(g_dec, g_mix, wa, wb):
    lib = _dev(g_dec)
    b = g_mix.shape[0]
    out = torch.empty_like(g_dec)
    _lib.check(lib.aesr_mix_bwd(g_dec.data_ptr(), g_mix.data_ptr(), wa.data_ptr(), wb.data_ptr(), out.data_ptr(), b,
                                g_mix[0].numel(), _stream(g_dec)), "mix_bwd")
    return out


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    lib = _dev(p)
    _lib.check(lib.aesr_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr),
                                  float(beta1), float(beta2), float(eps), float(weight_decay), int(step), _stream(p)),
               "adam_step")


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_dev: torch.Tensor,
                  lr_dev: Optional[torch.Tensor] = None):
    """Adam step whose step count (int32 [1]) and, optionally, learning rate (fp32 [1]) live in device memory:
    replayable inside a CUDA graph."""
    lib = _dev(p)
    assert step_dev.dtype == torch.int32 and step_dev.is_cuda
    assert lr_dev is None or (lr_dev.dtype == torch.float32 and lr_dev.is_cuda)
    _lib.check(lib.aesr_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr),
                                      float(beta1), float(beta2), float(eps), float(weight_decay), step_dev.data_ptr(),
                                      _ptr(lr_dev), _stream(p)), "adam_step_dev")


def vgg_conv1_fwd(img, w, b, shift3, scale3, normalize, dtype):
    lib = _dev(img)
    n, _, h, wd = img.shape
    out = torch.empty((n, h, wd, 64), dtype=dtype, device=img.device)
    with _timed("vgg_conv1", 2.0 * n * h * wd * 27 * 64):
        _lib.check(lib.aesr_vgg_conv1_fwd(img.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, h, wd,
                                          _F3(*shift3), _F3(*scale3), int(normalize), dt_code(dtype), _stream(img)),
                   "vgg_conv1_fwd")
    return out


def vgg_conv1_bwd(g, w, scale3, normalize, out_scale=1.0, out: Optional[torch.Tensor] = None):
    lib = _dev(g)
    n, h, wd, _ = g.shape
    dimg = out if out is not None else torch.empty((n, 1, h, wd), dtype=torch.float32, device=g.device)
    assert dimg.dtype == torch.float32 and dimg.is_contiguous() and dimg.numel() == n * h * wd
    with _timed("vgg_conv1"):
        _lib.check(lib.aesr_vgg_conv1_bwd(g.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, h, wd, _F3(*scale3),
                                          int(normalize), float(out_scale), _stream(g)), "vgg_conv1_bwd")
    return dimg


def maxpool_bwd(a, d_pooled, g_tap):
    lib = _dev(a)
    n, h, w, c = a.shape
    out = torch.empty((n, h, w, c), dtype=GRAD_DTYPE, device=a.device)
    with _timed("maxpool_bwd"):
        _lib.check(lib.aesr_maxpool_bwd(a.data_ptr(), _ptr(d_pooled), _ptr(g_tap), out.data_ptr(), n, h, w, c,
                                        dt_code(a.dtype), _stream(a)), "maxpool_bwd")
    return out


def lpips_head(o0, o1, lin, val: Optional[torch.Tensor], upstream: Optional[torch.Tensor] = None, want_grad=False):
    lib = _dev(o0)
    n, h, w, c = o0.shape
    g1 = torch.empty((n, h, w, c), dtype=GRAD_DTYPE, device=o0.device) if want_grad else None
    with _timed("lpips_head"):
        _lib.check(lib.aesr_lpips_head(o0.data_ptr(), o1.data_ptr(), lin.data_ptr(), _ptr(val), _ptr(upstream),
                                       _ptr(g1), n, h * w, c, dt_code(o0.dtype), _stream(o0)), "lpips_head")
    return g1
