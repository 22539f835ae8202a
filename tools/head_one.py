"""One launch shape of the fused decoder tail (dec.12 + head, 32 -> 4x32 @64x64), for ncu captures.

  AESR_HEAD_MMA=1 python tools/head_one.py [--n 3456] [--reps 3]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3456)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float16
x = torch.randn(a.n, 64, 64, 32, device=dev).to(dt)
b = torch.zeros(32, device=dev)
wp = ops.pack_conv3x3_weight_up2fold(torch.randn(32, 32, 3, 3, device=dev) * 0.05, dtype=dt)
hw9 = torch.randn(9, 32) * 0.1
out = torch.empty(a.n, 64, 64, 16, device=dev)
for _ in range(a.reps):
    ops.conv3x3_up2_head(x, wp, b, hw9, out=out)
torch.cuda.synchronize()
print("done")
