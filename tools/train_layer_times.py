"""Per-launch device times of one ACDC training step (CUDA events around every C-ABI call, warm, averaged).

  python tools/train_layer_times.py [--reps 5]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402
from oracle.make_golden import acdc_batch  # noqa: E402
from networks.net_config import NetworkConfig  # noqa: E402
from kwatsch.get_trainer import get_trainer_dynamic  # noqa: E402
from superresolution_aniso_mri_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
targs = dict(NetworkConfig("ae_combined", "ACDC").architecture)
targs.update(dataset="ACDC", model="ae_combined", ae_class="VanillaACAI", width=128, latent_width=32, latent=128,
             depth=32, lr=1e-5, weight_decay=0.0, epochs=10, device=str(dev), gpu_ids=[0],
             ex_loss_weight1=0.05, use_percept_loss=False, use_loss_annealing=False, get_masks=False,
             epoch_threshold=0, log_tensorboard=False, batch_size=12,
             _vgg_state=[t for pair in O.init_vgg(3) for t in pair])
torch.manual_seed(892372)
tr = get_trainer_dynamic(targs)
img, mid = (t.to(dev) for t in acdc_batch(0))
wa = torch.full((12,), 0.5, device=dev)


def step():
    tr.engine.step(img, mid, wa, wa, lpips=tr.percept_criterion, ex_loss_weight=0.05, lr=1e-5)


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
by = collections.Counter()
for (i, name, desc), (us, fl) in acc.items():
    us /= a.reps
    tot += us
    by[name] += us
    print("%3d %-12s %-40s %8.1f us %8.1f TFLOP/s" % (i, name, desc or "", us, fl / us / 1e6 if fl else 0.0))
print("total %.1f us" % tot, dict(by))
