"""The encoder stem at the ACDC per-step launch size, a few launches (ncu target).  python tools/stem_one.py"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200 import ops  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402

dev = torch.device("cuda:0")
args = O.default_args(128, 32)
args["device"] = "cuda:0"
m = VanillaACAI(args).eval()
m.load_state_dict(O.calibrated_state(O.default_args(128, 32)))
sp = m._stem()
xs = torch.rand(640, 1, 128, 128, device=dev)
for _ in range(3):
    ops.stem(xs, sp)
torch.cuda.synchronize()
print("ok")
