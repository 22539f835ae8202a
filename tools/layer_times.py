"""Per-launch device times of the inference pipelines (CUDA events around every C-ABI call, warm, averaged).

  python tools/layer_times.py [--enc 256] [--dec 252] [--reps 10]

Prints one line per launch: kernel family, shape, average us, algorithmic TFLOP/s.  Used to decide which layer to work on;
numbers for the record come from bench.py and ncu (profiles/).
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200 import ops  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--enc", type=int, default=256)
ap.add_argument("--dec", type=int, default=252)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--width", type=int, default=128)
ap.add_argument("--latent_width", type=int, default=32)
a = ap.parse_args()

args = O.default_args(a.width, a.latent_width)
args["device"] = "cuda:0"
model = VanillaACAI(args)
model.load_state_dict(O.calibrated_state(O.default_args(a.width, a.latent_width)))
model.eval()
dev = torch.device("cuda:0")
x = torch.rand(a.enc, 1, a.width, a.width, device=dev)
z = model.encode_eval(x)
P, K = a.dec // 6, 6
pa = (torch.arange(P, dtype=torch.int32, device=dev) + 1) % a.enc
pb = torch.arange(P, dtype=torch.int32, device=dev) % a.enc
wa = torch.linspace(0.1, 0.9, K, device=dev)
wb = 1 - wa


def step():
    model.encode_eval(x)
    lat = ops.lerp_pairs(z, pa, pb, wa, wb)
    model.decode_nhwc_eval(lat)


for _ in range(3):
    step()
acc = collections.OrderedDict()
for _ in range(a.reps):
    ops.TIMING = []
    step()
    torch.cuda.synchronize()
    for i, (name, e0, e1, fl, desc) in enumerate(ops.TIMING):
        r = acc.setdefault((i, name, desc), [0.0, fl])
        r[0] += e0.elapsed_time(e1) * 1e3
    ops.TIMING = None
tot = 0.0
for (i, name, desc), (us, fl) in acc.items():
    us /= a.reps
    tot += us
    print("%2d %-8s %-34s %8.1f us %8.1f TFLOP/s" % (i, name, desc, us, fl / us / 1e6 if fl else 0.0))
print("total %.1f us (fused=%s)" % (tot, model.fused_inference))
