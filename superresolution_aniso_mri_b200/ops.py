"""Tensor-level wrappers over the C-ABI (device pointers + current CUDA stream in, nothing else).

PyTorch is plumbing here: it owns the device memory and the stream; all arithmetic happens in libaesr_b200.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
OUT_SAME, OUT_AVGPOOL2, OUT_UP2, OUT_NCHW_F32, OUT_SAME_MAXPOOL2 = 0, 1, 2, 3, 4
MUL_NONE, MUL_LEAKY_GRAD, MUL_RELU_GRAD = 0, 1, 2
LEAKY_SLOPE = 0.01


def _dev(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("aesr_b200 operators run on a CUDA (sm_100a) device only; got a %s tensor -- there is no "
                           "CPU fallback" % t.device)
    return _lib.lib_for_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pack_conv3x3_weight(w: torch.Tensor, transpose_flip: bool = False) -> torch.Tensor:
    """fp32 [Cout,Cin,3,3] -> bf16 [9,Cout,Cin] (or [9,Cin,Cout] with mirrored taps for the data-gradient conv)."""
    lib = _dev(w)
    w = w.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    shape = (9, cin, cout) if transpose_flip else (9, cout, cin)
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.aesr_pack_conv3x3_weight(w.data_ptr(), out.data_ptr(), cout, cin, int(transpose_flip), _stream(w)),
               "pack_conv3x3_weight")
    return out


def conv_out_shape(n, h, w, cout, out_mode) -> Tuple[int, ...]:
    if out_mode == OUT_AVGPOOL2:
        return (n, h // 2, w // 2, cout)
    if out_mode == OUT_UP2:
        return (n, 2 * h, 2 * w, cout)
    if out_mode == OUT_NCHW_F32:
        return (n, cout, h, w)
    return (n, h, w, cout)


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE,
            slope: float = LEAKY_SLOPE, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
            out_mode: int = OUT_SAME, out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
            want_out2: bool = False, mul_src: Optional[torch.Tensor] = None, mul_mode: int = MUL_NONE,
            stats: Optional[torch.Tensor] = None):
    """x NHWC bf16 [N,H,W,Cin]; w_packed bf16 [9,Cout,Cin].  Returns out (and out2 when the mode produces one)."""
    lib = _dev(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, h, w, cin = x.shape
    cout = w_packed.shape[1]
    assert w_packed.shape == (9, cout, cin) and w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous()
    if out is None:
        out = torch.empty(conv_out_shape(n, h, w, cout, out_mode),
                          dtype=torch.float32 if out_mode == OUT_NCHW_F32 else torch.bfloat16, device=x.device)
    if out2 is None and (out_mode == OUT_SAME_MAXPOOL2 or (out_mode == OUT_NCHW_F32 and want_out2)):
        shp = (n, h // 2, w // 2, cout) if out_mode == OUT_SAME_MAXPOOL2 else (n, h, w, cout)
        out2 = torch.empty(shp, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.aesr_conv3x3_fwd(x.data_ptr(), w_packed.data_ptr(), _ptr(bias), _ptr(scale), _ptr(shift),
                                    out.data_ptr(), _ptr(out2), _ptr(mul_src), _ptr(stats), n, h, w, cin, cout,
                                    int(act), float(slope), int(out_mode), int(mul_mode), _stream(x)), "conv3x3_fwd")
    return (out, out2) if out2 is not None else out


def e0(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """enc.0: x fp32 [N,1,H,W] -> NHWC bf16 [N,H+2,W+2,C]."""
    lib = _dev(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 1
    n, _, h, wd = x.shape
    c = w.numel()
    out = torch.empty((n, h + 2, wd + 2, c), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.aesr_e0_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, h, wd, c, _stream(x)),
               "e0_fwd")
    return out


def head(a: torch.Tensor, w9c: torch.Tensor, bias: float, out: Optional[torch.Tensor] = None,
         out_image_stride: Optional[int] = None, sigmoid: bool = True,
         out_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dec.14 + sigmoid: NHWC bf16 [N,H,W,32] -> fp32 [N,1,H,W] (or image n -> out[out_index[n]])."""
    lib = _dev(a)
    n, h, w, c = a.shape
    if out is None:
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=a.device)
        out_image_stride = h * w
    _lib.check(lib.aesr_head_fwd(a.data_ptr(), w9c.data_ptr(), float(bias), out.data_ptr(), _ptr(out_index), n, h, w, c,
                                 int(out_image_stride), int(sigmoid), _stream(a)), "head_fwd")
    return out


def lerp_latents(z: torch.Tensor, ia: torch.Tensor, ib: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
                 want_nchw: bool = False):
    """out[m] = wa[m]*z[ia[m]] + wb[m]*z[ib[m]]; z fp32 NCHW -> NHWC bf16 [M,h,w,C] (+ fp32 NCHW z_mix)."""
    lib = _dev(z)
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 4
    _, c, h, w = z.shape
    m = ia.numel()
    out = torch.empty((m, h, w, c), dtype=torch.bfloat16, device=z.device)
    out_nchw = torch.empty((m, c, h, w), dtype=torch.float32, device=z.device) if want_nchw else None
    done = 0
    while done < m:                                  # gridDim.z limit
        cnt = min(m - done, 65535)
        _lib.check(lib.aesr_lerp_latents(z.data_ptr(), ia[done:].data_ptr(), ib[done:].data_ptr(),
                                         wa[done:].data_ptr(), wb[done:].data_ptr(), out[done:].data_ptr(),
                                         _ptr(out_nchw[done:]) if want_nchw else None, cnt, c, h * w, _stream(z)),
                   "lerp_latents")
        done += cnt
    return (out, out_nchw) if want_nchw else out


def place_slices(src: torch.Tensor, dst: torch.Tensor, out_index: Optional[torch.Tensor], clamp: bool = True) -> None:
    """dst[out_index[n]] = clamp(src[n], 0, 1) for fp32 images; src [N,HW...] contiguous, dst [*,HW...] contiguous."""
    lib = _dev(src)
    assert src.dtype == torch.float32 and dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    n = src.shape[0]
    hw = src[0].numel()
    done = 0
    while done < n:
        cnt = min(n - done, 65535)
        _lib.check(lib.aesr_place_slices(src[done:].data_ptr(), dst.data_ptr(),
                                         None if out_index is None else out_index[done:].data_ptr(), cnt, hw,
                                         int(clamp), _stream(src)), "place_slices")
        done += cnt
