"""A/B of the 32 -> 32 conv layers: conv3x3_fold_kernel (horizontal taps folded into N = 96) in its tile shapes against the
tap-by-tap halo kernel, at the per-step launch sizes of the ACDC inference pipeline and of the B = 12 training step.

  python tools/fold_sweep.py [--reps 10]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float16


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


LAYERS = [("dec.8 / dec.10", 64, 3456, 0), ("enc.3 pool", 130, 640, 1), ("train enc.3 (36 x 162^2)", 162, 36, 0),
          ("train dec.8 (36 x 80^2)", 80, 36, 0), ("oasis dec.10 (720 x 110^2)", 110, 720, 0)]
for name, hw, n, mode in LAYERS:
    x = torch.randn(n, hw, hw, 32, device=dev).to(dt)
    b = torch.zeros(32, device=dev)
    sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    wp = ops.pack_conv3x3_weight(torch.randn(32, 32, 3, 3, device=dev) * 0.05, dtype=dt)
    out = torch.empty(ops.conv_out_shape(n, hw, hw, 32, mode), dtype=dt, device=dev)
    fn = lambda: ops.conv3x3(x, wp, b, act=1, scale=sc, shift=sh, out_mode=mode, out=out)      # noqa: E731
    flops = 2.0 * n * hw * hw * 9 * 32 * 32
    bytes_ = x.numel() * 2 + out.numel() * 2
    print("== %s: 32->32 @%d n=%d mode=%d   HBM floor %.3f ms at 6.5 TB/s" % (name, hw, n, mode, bytes_ / 6.5e9))
    runs = [("halo (tap by tap)", {ops.TUNE_FOLD: 0}), ("fold T=1 nbuf=5", {}), ("fold T=1 nbuf=4", {ops.TUNE_CONV_NBUF: 4}),
            ("fold T=2 nbuf=2", {ops.TUNE_CONV_T: 2}), ("fold T=1 nbuf=5 stages=4", {ops.TUNE_CONV_STAGES: 4}),
            ("fold T=1 nbuf=5 stages=2", {ops.TUNE_CONV_STAGES: 2})]
    for label, kw in runs:
        for k in (ops.TUNE_FOLD, ops.TUNE_CONV_T, ops.TUNE_CONV_NBUF, ops.TUNE_CONV_STAGES):
            ops.set_tuning(k, kw.get(k, 1 if k == ops.TUNE_FOLD else 0))
        ms = timed(fn)
        print("  %-26s %7.3f ms  %7.1f TFLOP/s  %6.0f GB/s" % (label, ms, flops / ms / 1e9, bytes_ / ms / 1e6))
    for k in (ops.TUNE_FOLD, ops.TUNE_CONV_T, ops.TUNE_CONV_NBUF, ops.TUNE_CONV_STAGES):
        ops.set_tuning(k, 0)
