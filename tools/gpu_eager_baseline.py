"""The "same-box GPU bar" of SURVEY.md section 8(d): the reference's algorithm for the path (oracle port =
torch functional fp32 ops, cuDNN / cuBLAS underneath, TF32 off) run on the B200 itself, next to this repo's kernels.

  python tools/gpu_eager_baseline.py [--volumes 16]

Two variants: (a) as the reference is written (generate_hr_volumes.create_super_volume: volume by volume, both
neighbours re-encoded for every alpha, torch.cat chain); (b) minimal-work batching of the same fp32 ops (every slice
encoded once, all blends decoded as one batch) -- what a careful PyTorch user would get without custom kernels.
Diagnostic only (tools/): the oracle is test infrastructure, never part of the product path.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--volumes", type=int, default=16)
a = ap.parse_args()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
args = O.default_args(128, 32)
state = {k: v.to(dev) for k, v in O.calibrated_state(args).items()}
ar = O.alpha_range_for(6)
vols = torch.rand(a.volumes, 10, 1, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
n_synth = a.volumes * 9 * 6


def as_written():
    for v in vols:
        O.create_super_volume(state, args, v, ar, use_original=True)


def minimal_work():
    with torch.no_grad():
        flat = vols.view(-1, 1, 128, 128)
        z = O.encode(state, args, flat).view(a.volumes, 10, 128, 32, 32)
        hi, lo = O.interp_weights(ar)
        hi_t = torch.from_numpy(hi).to(dev).view(1, 1, 6, 1, 1, 1)
        lo_t = torch.from_numpy(lo).to(dev).view(1, 1, 6, 1, 1, 1)
        mix = hi_t * z[:, 1:, None] + lo_t * z[:, :-1, None]            # [V, 9, 6, 128, 32, 32]
        out = O.decode(state, args, mix.view(-1, 128, 32, 32))
        return torch.clamp(out, 0, 1)


for name, fn in (("as written (re-encode per alpha, per volume)", as_written), ("minimal work, batched fp32 eager", minimal_work)):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-46s %8.1f slices/s  (%d volumes, %.1f ms)" % (name, n_synth / dt, a.volumes, dt * 1e3))
