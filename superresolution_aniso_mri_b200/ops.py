"""Tensor-level wrappers over the C-ABI (device pointers + current CUDA stream in, nothing else).

PyTorch is plumbing here: it owns the device memory and the stream; all arithmetic happens in libaesr_b200.so.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib

ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
OUT_SAME, OUT_AVGPOOL2, OUT_UP2, OUT_NCHW_F32, OUT_SAME_MAXPOOL2, OUT_SHUFFLE2 = 0, 1, 2, 3, 4, 5
OUT_SAME_F32 = 7      # NHWC fp32, un-rounded accumulators (6 = the fused head, see conv3x3_up2_head)
MUL_NONE, MUL_LEAKY_GRAD, MUL_RELU_GRAD = 0, 1, 2
ALGO_AUTO, ALGO_HALO, ALGO_STREAM = 0, 1, 2
DT_BF16, DT_FP16 = 0, 1
LEAKY_SLOPE = 0.01

# 16-bit storage format of internal activations / packed filters.  fp16 (10-bit mantissa, fp32 accumulate) is the
# default for inference: 8x smaller rounding error than bf16 at the same tensor-core rate; BatchNorm keeps activations
# far inside the fp16 range.  AESR_ACT_DTYPE=bf16 switches globally.
DEFAULT_DTYPE = torch.bfloat16 if os.environ.get("AESR_ACT_DTYPE", "fp16").lower() == "bf16" else torch.float16


# Per-launch device timing for bench.py's roofline pass: when a list, every wrapper appends
# (kernel family, start event, end event, algorithmic FLOPs).  None (default) = no events, no overhead.
TIMING = None


class _timed:
    def __init__(self, name: str, flops: float = 0.0, desc: str = ""):
        self.name, self.flops, self.desc = name, flops, desc

    def __enter__(self):
        if TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if TIMING is not None:
            self.e1.record()
            TIMING.append((self.name, self.e0, self.e1, self.flops, self.desc))
        return False


def dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return DT_FP16
    if dtype == torch.bfloat16:
        return DT_BF16
    raise RuntimeError("aesr_b200: internal activations are fp16 or bf16, got %s" % dtype)


def _dev(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("aesr_b200 operators run on a CUDA (sm_100a) device only; got a %s tensor -- there is no "
                           "CPU fallback" % t.device)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        raise RuntimeError("aesr_b200: tensor on cuda:%d but the current device is cuda:%d -- kernels launch on the current "
                           "device; wrap the call in `with torch.cuda.device(%d):` (one process per GPU is the supported "
                           "layout)" % (idx, torch.cuda.current_device(), idx))
    return _lib.lib_for_device(idx)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pack_conv3x3_weight(w: torch.Tensor, transpose_flip: bool = False, dtype: Optional[torch.dtype] = None
                        ) -> torch.Tensor:
    """fp32 [Cout,Cin,3,3] -> 16-bit [9,Cout,Cin] (or [9,Cin,Cout] with mirrored taps for the data-gradient conv)."""
    lib = _dev(w)
    dtype = dtype or DEFAULT_DTYPE
    w = w.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    shape = (9, cin, cout) if transpose_flip else (9, cout, cin)
    out = torch.empty(shape, dtype=dtype, device=w.device)
    _lib.check(lib.aesr_pack_conv3x3_weight(w.data_ptr(), out.data_ptr(), cout, cin, int(transpose_flip),
                                            dt_code(dtype), _stream(w)), "pack_conv3x3_weight")
    return out


def pack_conv3x3_weight_batch(src_base: torch.Tensor, dst_base: torch.Tensor, jobs_dev: torch.Tensor, max_elems: int) -> None:
    """All filters of a model in one launch: ``jobs_dev`` int64 [n,6] on the device = (offset in src_base, offset in
    dst_base, Cout, Cin, transpose_flip, 0) per filter; ``src_base`` fp32 flat parameters, ``dst_base`` 16-bit flat."""
    lib = _dev(src_base)
    assert src_base.dtype == torch.float32 and jobs_dev.dtype == torch.int64 and jobs_dev.is_cuda and jobs_dev.is_contiguous()
    _lib.check(lib.aesr_pack_conv3x3_weight_batch(src_base.data_ptr(), dst_base.data_ptr(), jobs_dev.data_ptr(),
                                                  jobs_dev.shape[0], int(max_elems), dt_code(dst_base.dtype),
                                                  _stream(src_base)), "pack_conv3x3_weight_batch")


def pack_conv3x3_weight_up2fold(w: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """fp32 [Cout,Cin,3,3] of a conv that FOLLOWS nn.Upsample(2) -> 16-bit [9, 4*Cout, Cin]: the same conv expressed on
    the low-resolution input, one block of Cout rows per output phase (2y+a, 2x+b) (conv3x3 out_mode OUT_SHUFFLE2)."""
    lib = _dev(w)
    dtype = dtype or DEFAULT_DTYPE
    w = w.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    out = torch.empty((9, 4 * cout, cin), dtype=dtype, device=w.device)
    _lib.check(lib.aesr_pack_conv3x3_weight_up2fold(w.data_ptr(), out.data_ptr(), cout, cin, dt_code(dtype), _stream(w)),
               "pack_conv3x3_weight_up2fold")
    return out


def conv_out_shape(n, h, w, cout, out_mode) -> Tuple[int, ...]:
    if out_mode == OUT_SHUFFLE2:
        return (n, 2 * h, 2 * w, cout // 4)
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
            stats: Optional[torch.Tensor] = None, algo: int = ALGO_AUTO, stats_split: int = 0):
    """x NHWC 16-bit [N,H,W,Cin]; w_packed same dtype [9,Cout,Cin].  Returns out (and out2 when the mode has one).
    ``stats_split``: images >= it accumulate their BatchNorm sums into the second [2*Cout] block of ``stats``."""
    lib = _dev(x)
    assert x.is_contiguous() and x.dim() == 4 and x.dtype == w_packed.dtype
    n, h, w, cin = x.shape
    cout = w_packed.shape[1]
    assert w_packed.shape == (9, cout, cin) and w_packed.is_contiguous()
    if out is None:
        out = torch.empty(conv_out_shape(n, h, w, cout, out_mode),
                          dtype=torch.float32 if out_mode in (OUT_NCHW_F32, OUT_SAME_F32) else x.dtype,
                          device=x.device)
    if out2 is None and (out_mode == OUT_SAME_MAXPOOL2 or (out_mode == OUT_NCHW_F32 and want_out2)):
        shp = (n, h // 2, w // 2, cout) if out_mode == OUT_SAME_MAXPOOL2 else (n, h, w, cout)
        out2 = torch.empty(shp, dtype=x.dtype, device=x.device)
    with _timed("conv3x3", 2.0 * n * h * w * 9 * cin * cout, "%d->%d n%d %dx%d mode%d" % (cin, cout, n, h, w, out_mode)):
        _lib.check(lib.aesr_conv3x3_fwd(x.data_ptr(), w_packed.data_ptr(), _ptr(bias), _ptr(scale), _ptr(shift),
                                        out.data_ptr(), _ptr(out2), _ptr(mul_src), _ptr(stats), n, h, w, cin, cout,
                                        int(act), float(slope), int(out_mode), int(mul_mode), dt_code(x.dtype),
                                        int(algo), int(stats_split), _stream(x)), "conv3x3_fwd")
    return (out, out2) if out2 is not None else out


# Decoder head conv inside the fused tail: False (default) = CUDA cores, fp32 activations and filter; True = tensor cores
# (activations written back to tensor memory as a 16-bit A operand).  Measured equal within a few % on B200 (the tensor-core
# variant is bound by the latency of its two-phase epilogue), so the more exact one is the default.
HEAD_ON_TENSOR_CORES = os.environ.get("AESR_HEAD_TC", "0") != "0"


TUNE_CONV_DEBUG, TUNE_CONV_T, TUNE_CONV_NBUF, TUNE_CONV_STAGES, TUNE_HEAD_MMA = range(5)
TUNE_FOLD = 9


def set_tuning(key: int, value: int) -> None:
    """``aesr_set_tuning``: kernel-variant / profiling knobs of the conv kernel (include/aesr_b200.h), 0 = automatic."""
    _lib.check(_lib.load().aesr_set_tuning(int(key), int(value)), "set_tuning")


def pack_head_w16(head_w9c: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """fp32 [9,32] head filter (any device) -> 16-bit [16,32] on its device, rows 9..15 zero (GEMM N padded to 16)."""
    dtype = dtype or DEFAULT_DTYPE
    out = torch.zeros((16, 32), dtype=dtype, device=head_w9c.device)
    out[:9] = head_w9c.to(dtype)
    return out


def conv3x3_up2_head(x: torch.Tensor, w_folded: torch.Tensor, bias: torch.Tensor, head_w9c: torch.Tensor,
                     act: int = ACT_LEAKY, slope: float = LEAKY_SLOPE, out: Optional[torch.Tensor] = None,
                     algo: int = ALGO_AUTO, head_w16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Upsample(2) -> Conv2d(Cin,32,3,p=1)+act -> Conv2d(32,1,3,p=1) partial sums: x NHWC 16-bit [N,H,W,Cin] (low-res)
    -> fp32 [N,H,W,16] patches for ``head_gather``.  ``head_w9c``: fp32 [9,32] on the HOST; ``head_w16``: the same
    filter as 16-bit [16,32] on the device (``pack_head_w16``) -> the head conv runs on the tensor cores."""
    lib = _dev(x)
    assert x.is_contiguous() and x.dim() == 4 and x.dtype == w_folded.dtype
    n, h, w, cin = x.shape
    assert w_folded.shape == (9, 128, cin) and w_folded.is_contiguous() and head_w9c.shape == (9, 32)
    # the head filter travels as a kernel parameter (constant bank): it must be HOST memory
    assert not head_w9c.is_cuda and head_w9c.dtype == torch.float32 and head_w9c.is_contiguous()
    if head_w16 is not None:
        assert head_w16.is_cuda and head_w16.shape == (16, 32) and head_w16.dtype == x.dtype and head_w16.is_contiguous()
    if out is None:
        out = torch.empty((n, h, w, 16), dtype=torch.float32, device=x.device)
    with _timed("conv3x3", 2.0 * n * h * w * 9 * cin * 128 + 2.0 * n * 4 * h * w * 9 * 32,
                "%d->128+head n%d %dx%d" % (cin, n, h, w)):
        _lib.check(lib.aesr_conv3x3_up2_head_fwd(x.data_ptr(), w_folded.data_ptr(), bias.data_ptr(), head_w9c.data_ptr(),
                                                 _ptr(head_w16), out.data_ptr(), n, h, w, cin, int(act), float(slope),
                                                 dt_code(x.dtype), int(algo), _stream(x)), "conv3x3_up2_head_fwd")
    return out


def head_gather(partial: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None,
                out_image_stride: Optional[int] = None, sigmoid: bool = True,
                out_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """partial fp32 [N,h,w,16] -> fp32 images [N,1,2h,2w] (or image n -> out[out_index[n]])."""
    lib = _dev(partial)
    n, h, w, _ = partial.shape
    if out is None:
        out = torch.empty((n, 1, 2 * h, 2 * w), dtype=torch.float32, device=partial.device)
        out_image_stride = 4 * h * w
    with _timed("head"):
        _lib.check(lib.aesr_head_gather(partial.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(out_index), n, h, w,
                                        int(out_image_stride), int(sigmoid), _stream(partial)), "head_gather")
    return out


def stem_fold(w0: torch.Tensor, b0: torch.Tensor, w1: torch.Tensor):
    """enc.0 (1x1, 1->C) composed with enc.1 (3x3, C->C): effective single-channel 3x3 filter and bias taps [9,C]."""
    lib = _dev(w1)
    c = w1.shape[0]
    weff = torch.empty((9, c), dtype=torch.float32, device=w1.device)
    beff = torch.empty((9, c), dtype=torch.float32, device=w1.device)
    _lib.check(lib.aesr_stem_fold(w0.detach().reshape(-1).contiguous().data_ptr(), b0.detach().contiguous().data_ptr(),
                                  w1.detach().contiguous().data_ptr(), weff.data_ptr(), beff.data_ptr(), c, _stream(w1)),
               "stem_fold")
    return weff, beff


def stem_host_params(weff: torch.Tensor, beff: torch.Tensor, b1: torch.Tensor) -> torch.Tensor:
    """[weff | beff | b1] as one fp32 HOST tensor (608 floats): travels to ``stem`` as kernel parameters."""
    return torch.cat([t.detach().reshape(-1).float().cpu() for t in (weff, beff, b1)]).contiguous()


def stem(x: torch.Tensor, params_host: torch.Tensor, slope: float = LEAKY_SLOPE,
         dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """enc.0 + enc.1 + LeakyReLU: x fp32 [N,1,H,W] -> NHWC 16-bit [N,H+2,W+2,32]."""
    lib = _dev(x)
    dtype = dtype or DEFAULT_DTYPE
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 1
    assert not params_host.is_cuda and params_host.numel() == 608 and params_host.dtype == torch.float32
    n, _, h, wd = x.shape
    c = 32
    out = torch.empty((n, h + 2, wd + 2, c), dtype=dtype, device=x.device)
    with _timed("stem", 2.0 * n * (h + 2) * (wd + 2) * c * (1 + 9 * c)):
        _lib.check(lib.aesr_stem_fwd(x.data_ptr(), params_host.data_ptr(), out.data_ptr(), n, h, wd, c, float(slope),
                                     dt_code(dtype), _stream(x)), "stem_fwd")
    return out


def e0(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """enc.0: x fp32 [N,1,H,W] -> NHWC 16-bit [N,H+2,W+2,C]."""
    lib = _dev(x)
    dtype = dtype or DEFAULT_DTYPE
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 1
    n, _, h, wd = x.shape
    c = w.numel()
    out = torch.empty((n, h + 2, wd + 2, c), dtype=dtype, device=x.device)
    with _timed("e0", 2.0 * n * (h + 2) * (wd + 2) * c):
        _lib.check(lib.aesr_e0_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, h, wd, c,
                                   dt_code(dtype), _stream(x)), "e0_fwd")
    return out


def head(a: torch.Tensor, w9c: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None,
         out_image_stride: Optional[int] = None, sigmoid: bool = True,
         out_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dec.14 + sigmoid: NHWC 16-bit [N,H,W,32] -> fp32 [N,1,H,W] (or image n -> out[out_index[n]])."""
    lib = _dev(a)
    n, h, w, c = a.shape
    if out is None:
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=a.device)
        out_image_stride = h * w
    with _timed("head", 2.0 * n * h * w * 9 * c):
        _lib.check(lib.aesr_head_fwd(a.data_ptr(), w9c.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(out_index), n, h,
                                     w, c, int(out_image_stride), int(sigmoid), dt_code(a.dtype), _stream(a)),
                   "head_fwd")
    return out


def lerp_latents(z: torch.Tensor, ia: torch.Tensor, ib: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
                 want_nchw: bool = False, dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None):
    """out[m] = wa[m]*z[ia[m]] + wb[m]*z[ib[m]]; z fp32 NCHW -> NHWC 16-bit [M,h,w,C] (+ fp32 NCHW z_mix)."""
    lib = _dev(z)
    dtype = dtype or (out.dtype if out is not None else DEFAULT_DTYPE)
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 4
    _, c, h, w = z.shape
    m = ia.numel()
    if out is None:
        out = torch.empty((m, h, w, c), dtype=dtype, device=z.device)
    assert out.shape == (m, h, w, c) and out.dtype == dtype and out.is_contiguous()
    out_nchw = torch.empty((m, c, h, w), dtype=torch.float32, device=z.device) if want_nchw else None
    done = 0
    tm = _timed("lerp").__enter__()
    while done < m:                                  # gridDim.z limit
        cnt = min(m - done, 65535)
        _lib.check(lib.aesr_lerp_latents(z.data_ptr(), ia[done:].data_ptr(), ib[done:].data_ptr(),
                                         wa[done:].data_ptr(), wb[done:].data_ptr(), out[done:].data_ptr(),
                                         out_nchw[done:].data_ptr() if want_nchw else None, cnt, c, h * w,
                                         dt_code(dtype), _stream(z)), "lerp_latents")
        done += cnt
    return (out, out_nchw) if want_nchw else out


def lerp_pairs(z: torch.Tensor, pa: torch.Tensor, pb: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
               dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out[p*K + k] = wa[k]*z[pa[p]] + wb[k]*z[pb[p]]; z fp32 NCHW -> NHWC 16-bit [P*K,h,w,C] (latents read once)."""
    lib = _dev(z)
    dtype = dtype or DEFAULT_DTYPE
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 4
    _, c, h, w = z.shape
    p, k = pa.numel(), wa.numel()
    out = torch.empty((p * k, h, w, c), dtype=dtype, device=z.device)
    with _timed("lerp"):
        done = 0
        while done < p:
            cnt = min(p - done, 65535)
            _lib.check(lib.aesr_lerp_pairs(z.data_ptr(), pa[done:].data_ptr(), pb[done:].data_ptr(), wa.data_ptr(),
                                           wb.data_ptr(), out[done * k:].data_ptr(), cnt, k, c, h * w, dt_code(dtype),
                                           _stream(z)), "lerp_pairs")
            done += cnt
    return out


def lerp_pairs_act(pre: torch.Tensor, pa: torch.Tensor, pb: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor,
                   bias: Optional[torch.Tensor], slope: float = LEAKY_SLOPE, dtype: Optional[torch.dtype] = None
                   ) -> torch.Tensor:
    """out[p*K + k] = LeakyReLU(wa[k]*pre[pa[p]] + wb[k]*pre[pb[p]] + bias); pre fp32 NHWC [*,h,w,C] (the decoder's
    first conv applied to the latents, OUT_SAME_F32) -> NHWC 16-bit [P*K,h,w,C]."""
    lib = _dev(pre)
    dtype = dtype or DEFAULT_DTYPE
    assert pre.dtype == torch.float32 and pre.is_contiguous() and pre.dim() == 4
    _, h, w, c = pre.shape
    p, k = pa.numel(), wa.numel()
    out = torch.empty((p * k, h, w, c), dtype=dtype, device=pre.device)
    with _timed("lerp"):
        done = 0
        while done < p:
            cnt = min(p - done, 65535)
            _lib.check(lib.aesr_lerp_pairs_act(pre.data_ptr(), pa[done:].data_ptr(), pb[done:].data_ptr(),
                                               wa.data_ptr(), wb.data_ptr(), _ptr(bias), out[done * k:].data_ptr(),
                                               cnt, k, c, h * w, float(slope), dt_code(dtype), _stream(pre)),
                       "lerp_pairs_act")
            done += cnt
    return out


def copy_rows_async(dst: torch.Tensor, src: torch.Tensor, dst_offset: int, src_offset: int, outer: int,
                    dst_outer_stride: int, src_outer_stride: int, rows: int, dpitch: int, spitch: int, width: int,
                    stream: torch.cuda.Stream) -> None:
    """``outer`` strided 2-D copies between a PINNED host tensor and a device tensor (either direction, taken from
    which of the two is on the device); offsets / strides / pitches / width in BYTES.  Stream-ordered on ``stream``.
    (A download done by a few thread blocks storing straight into the pinned buffer reaches the same 52 GB/s alone but
    starves behind the persistent conv CTAs inside the pipeline: 7.0 vs 5.1 ms per step,
    profiles/r02n_e2e_sm_download_probe.txt -- the copy engine stays.)"""
    to_host = src.is_cuda
    lib = _dev(src if to_host else dst)
    host = dst if to_host else src
    assert host.is_pinned() and dst.is_contiguous() and src.is_contiguous()
    _lib.check(lib.aesr_copy_rows_async(dst.data_ptr() + dst_offset, dst_outer_stride, dpitch, src.data_ptr() + src_offset,
                                        src_outer_stride, spitch, width, rows, outer, int(to_host), stream.cuda_stream),
               "copy_rows_async")


def place_slices(src: torch.Tensor, dst: torch.Tensor, out_index: Optional[torch.Tensor], clamp: bool = True) -> None:
    """dst[out_index[n]] = clamp(src[n], 0, 1) for fp32 images; src [N,HW...] contiguous, dst [*,HW...] contiguous."""
    lib = _dev(src)
    assert src.dtype == torch.float32 and dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    n = src.shape[0]
    hw = src[0].numel()
    done = 0
    while done < n:
        cnt = min(n - done, 65535)
        _lib.check(lib.aesr_place_slices(src[done:].data_ptr(), dst.data_ptr() if out_index is not None else dst[done:].data_ptr(),
                                         None if out_index is None else out_index[done:].data_ptr(), cnt, hw,
                                         int(clamp), _stream(src)), "place_slices")
        done += cnt
