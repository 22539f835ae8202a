"""Encoder stem with a cold L2 (256 MB flush before every launch), CUDA-core vs warp-MMA version.

  python tools/stem_cold.py

Measured (r03b): 0.244 / 0.208 ms -- the same as with a warm L2 and for L2 prefetch distances 2..16 images: the input is
6 % of the traffic and its latency is covered.
"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from oracle import aesr_oracle as O
from superresolution_aniso_mri_b200 import ops
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI
dev = torch.device("cuda:0")
args = O.default_args(128, 32); args["device"] = "cuda:0"
m = VanillaACAI(args).eval()
sp = m._stem()
xs = torch.rand(640, 1, 128, 128, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def cold(reps=8):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.stem(xs, sp); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for v in (1, 0):
    ops.set_tuning(5, v)
    print("variant %d (1 = CUDA cores, 0 = warp MMA) cold L2: %.3f ms" % (v, cold()))
