"""One eager ACDC training step under the CUDA profiler range (for `ncu --profile-from-start off`).

  ncu --profile-from-start off --set full --clock-control none -o gpurun_out/prof_train python tools/train_ncu.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402
from oracle.make_golden import acdc_batch  # noqa: E402
from networks.net_config import NetworkConfig  # noqa: E402
from kwatsch.get_trainer import get_trainer_dynamic  # noqa: E402

dev = torch.device("cuda:0")
targs = dict(NetworkConfig("ae_combined", "ACDC").architecture)
targs.update(dataset="ACDC", model="ae_combined", ae_class="VanillaACAI", width=128, latent_width=32, latent=128,
             depth=32, lr=1e-5, weight_decay=0.0, epochs=10, device=str(dev), gpu_ids=[0],
             ex_loss_weight1=0.05, use_percept_loss=False, use_loss_annealing=False, get_masks=False,
             epoch_threshold=0, log_tensorboard=False, batch_size=12,
             _vgg_state=[t for pair in O.init_vgg(3) for t in pair])
torch.manual_seed(892372)
tr = get_trainer_dynamic(targs)
tr.engine.use_graph = False
img, mid = (t.to(dev) for t in acdc_batch(0))
wa = torch.full((12,), 0.5, device=dev)
for _ in range(3):
    tr.engine.step(img, mid, wa, wa, lpips=tr.percept_criterion, ex_loss_weight=0.05, lr=1e-5)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.engine.step(img, mid, wa, wa, lpips=tr.percept_criterion, ex_loss_weight=0.05, lr=1e-5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
