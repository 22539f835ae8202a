"""Stage isolation of the fused decoder tail (dec.12 + head, 32 -> 4x32 @64x64, 3456 slices) for both head variants.

  python tools/head_sweep.py [--reps 5]

Stage mask (aesr_set_tuning key 0): 2 no activation TMA loads | 4 no stores | 8 no TMEM reads | 32 no MMAs.
Also times the encoder stem on the CUDA cores against the warp-MMA version.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--stem-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float16


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


n, hw, cin = 3456, 64, 32
x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
b = torch.zeros(32, device=dev)
wp = ops.pack_conv3x3_weight_up2fold(torch.randn(32, cin, 3, 3, device=dev) * 0.05, dtype=dt)
hw9 = torch.randn(9, 32) * 0.1
out = torch.empty(n, hw, hw, 16, device=dev)
fn = lambda: ops.conv3x3_up2_head(x, wp, b, hw9, out=out)      # noqa: E731
for variant in (() if a.stem_only else (0, 1)):
    ops.set_tuning(ops.TUNE_HEAD_MMA, variant)
    print("== dec.12+head, head on %s" % ("warp MMA" if variant else "CUDA cores"))
    for dbg in (0, 4, 8, 12, 32, 2, 34, 40, 44, 46):
        ops.set_tuning(ops.TUNE_CONV_DEBUG, dbg)
        print("  dbg=%-3d %.3f ms" % (dbg, timed(fn)))
    ops.set_tuning(ops.TUNE_CONV_DEBUG, 0)
    for T, nb in ((2, 2), (1, 2), (1, 4)):
        ops.set_tuning(ops.TUNE_CONV_T, T)
        ops.set_tuning(ops.TUNE_CONV_NBUF, nb)
        print("  T=%d nbuf=%d %.3f ms" % (T, nb, timed(fn)))
    ops.set_tuning(ops.TUNE_CONV_T, 0)
    ops.set_tuning(ops.TUNE_CONV_NBUF, 0)
ops.set_tuning(ops.TUNE_HEAD_MMA, 0)
del x, out

from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402
args = O.default_args(128, 32)
args["device"] = "cuda:0"
m = VanillaACAI(args).eval()
sp = m._stem()
xs = torch.rand(640, 1, 128, 128, device=dev)
for variant in (1, 0, 2, 1, 0, 2):
    ops.set_tuning(5, variant)
    ms = timed(lambda: ops.stem(xs, sp))
    print("stem n=640 (%s): %.3f ms  %.0f GB/s (algorithmic 4 B/px in + 64 B/px out)" % (
        ("warp MMA, 3 tf32 terms", "CUDA cores", "warp MMA, 1 tf32 term")[variant], ms,
        640 * (128 * 128 * 4 + 130 * 130 * 64) / ms / 1e6))
ops.set_tuning(5, 1)
r0 = ops.stem(xs[:8], sp).float()
ops.set_tuning(5, 0)
r1 = ops.stem(xs[:8], sp).float()
print("stem variants max-abs diff %.3e (max |value| %.3f)" % ((r0 - r1).abs().max().item(), r0.abs().max().item()))
