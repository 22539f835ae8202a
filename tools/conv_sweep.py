"""Stage-isolation / tuning sweep of the halo conv kernel in ONE process (aesr_set_tuning instead of env variables).

  python tools/conv_sweep.py [--reps 5] [--quick]

For every layer shape of the ACDC inference pipeline at its per-step launch size: time with the automatic
configuration, with the stage mask (2 no TMA loads | 4 no stores | 8 no TMEM loads | 32 no MMAs | 64 no epilogue
math), and with forced (T, nbuf) shapes.  Prints one line per run; the MMA-issue bound of each layer is printed next
to it (cycles per MMA measured by tools/umma_rate.py: N=32 40.3, N=64 48.2, N=128 64.2 at 1965 MHz on 148 SMs).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float16
lib = _lib.lib_for_device(0)
CYC = {32: 40.3, 64: 48.2, 128: 64.2, 256: 128.2}


def tune(debug=0, T=0, nbuf=0, stages=0):
    for k, v in enumerate((debug, T, nbuf, stages)):
        _lib.check(lib.aesr_set_tuning(k, v))


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


# name, cin, cout (GEMM N incl. phases), hw, n, mode ("head" = fused decoder tail)
LAYERS = [("dec.12+head TC", 32, 128, 64, 3456, "head"), ("dec.12+head", 32, 128, 64, 3456, "head"), ("dec.6 shuffle", 64, 128, 32, 3456, 5),
          ("dec.8", 32, 32, 64, 3456, 0), ("dec.2", 64, 64, 32, 3456, 0), ("dec.0 pre f32", 128, 64, 32, 640, 7),
          ("enc.3 pool", 32, 32, 130, 640, 1), ("enc.7", 32, 64, 65, 640, 0), ("enc.9 pool", 64, 64, 65, 640, 1),
          ("enc.13", 64, 128, 32, 640, 0), ("enc.15", 128, 128, 32, 640, 0)]
for name, cin, cout, hw, n, mode in LAYERS:
    x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
    b = torch.zeros(cout if mode != "head" else 32, device=dev)
    if mode == "head":
        wp = ops.pack_conv3x3_weight_up2fold(torch.randn(32, cin, 3, 3, device=dev) * 0.05, dtype=dt)
        hw9 = torch.randn(9, 32) * 0.1
        out = torch.empty(n, hw, hw, 16, device=dev)
        hw16 = ops.pack_head_w16(hw9.to(dev), dtype=dt) if name.endswith("TC") else None
        fn = lambda: ops.conv3x3_up2_head(x, wp, b, hw9, out=out, head_w16=hw16)      # noqa: E731
    else:
        wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, dtype=dt)
        out = torch.empty(ops.conv_out_shape(n, hw, hw, cout, mode), dtype=torch.float32 if mode == 7 else dt, device=dev)
        fn = lambda: ops.conv3x3(x, wp, b, act=1, out_mode=mode, out=out)      # noqa: E731
    he = (hw // 2) * 2 if mode == 1 else hw
    tiles = n * ((he + 7) // 8) * ((he + 15) // 16)
    bn = cout if cin * cout * 18 <= 160 * 1024 else cout // 2
    bound_ms = tiles * (cout // bn) * 9 * (cin // 16) * CYC[bn] / 148 / 1.965e6
    flops = 2.0 * n * hw * hw * 9 * cin * cout
    print("== %s: %d->%d @%d n=%d mode=%s  MMA-issue bound %.3f ms" % (name, cin, cout, hw, n, mode, bound_ms))
    runs = [("auto", dict())]
    if not a.quick:
        runs += [("dbg=%d" % d, dict(debug=d)) for d in ((32, 256) if mode == "head" else (4, 76, 78, 78 + 256, 256, 32, 2, 110))]
    runs += [("T=%d nbuf=%d" % (T, nb), dict(T=T, nbuf=nb)) for T, nb in ((4, 2), (4, 4), (2, 2), (2, 4), (1, 2), (1, 4))
             if T * nb * bn <= 512]
    if not a.quick:
        runs += [("stages=%d" % s_, dict(stages=s_)) for s_ in (2, 3)]
    for label, kw in runs:
        tune(**kw)
        try:
            ms = timed(fn)
            print("  %-14s %.3f ms  %7.1f TFLOP/s  %.0f%% of bound" % (label, ms, flops / ms / 1e9, 100 * bound_ms / ms))
        except RuntimeError as ex:
            print("  %-14s failed: %s" % (label, str(ex)[:100]))
    tune()
    del x, out
# memory-bound neighbours
x = torch.rand(640, 1, 128, 128, device=dev)
from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402
args = O.default_args(128, 32)
args["device"] = "cuda:0"
m = VanillaACAI(args).eval()
sp = m._stem()
ms = timed(lambda: ops.stem(x, sp))
print("stem n=640: %.3f ms  %.0f GB/s (algorithmic 4 B/px in + 64 B/px out)" % (ms, 640 * (128 * 128 * 4 + 130 * 130 * 64) / ms / 1e6))
part = torch.randn(3456, 64, 64, 16, device=dev)
outv = torch.empty(3456, 128, 128, device=dev)
bias = torch.zeros(1, device=dev)
ms = timed(lambda: ops.head_gather(part, bias, out=outv, out_image_stride=128 * 128))
print("head_gather n=3456: %.3f ms  %.0f GB/s (algorithmic 16 B/px in + 4 B/px out)" % (ms, 3456 * 128 * 128 * 20 / ms / 1e6))
