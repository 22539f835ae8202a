"""Is the training step launch-bound on the host?  Times (a) host enqueue time of one step without synchronising,
(b) device time per step, (c) the same step captured in a CUDA graph and replayed.

  python tools/train_graph_probe.py [steps]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402
from oracle.make_golden import acdc_batch  # noqa: E402
from networks.net_config import NetworkConfig  # noqa: E402
from kwatsch.get_trainer import get_trainer_dynamic  # noqa: E402
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
targs = dict(NetworkConfig("ae_combined", "ACDC").architecture)
targs.update(dataset="ACDC", model="ae_combined", ae_class="VanillaACAI", width=128, latent_width=32, latent=128,
             depth=32, lr=1e-5, weight_decay=0.0, epochs=10, device=str(dev), gpu_ids=[0], ex_loss_weight1=0.05,
             use_percept_loss=False, use_loss_annealing=False, get_masks=False, epoch_threshold=0,
             log_tensorboard=False, batch_size=12, _vgg_state=[t for pair in O.init_vgg(3) for t in pair])
torch.manual_seed(892372)
tr = get_trainer_dynamic(targs)
eng, lp = tr.engine, tr.percept_criterion
img, mid = [t.to(dev) for t in acdc_batch(0)]
wa = torch.full((12,), 0.5, device=dev)


def step():
    return eng.step(img, mid, wa, wa, lpips=lp, ex_loss_weight=0.05, lr=1e-5)


for _ in range(3):
    step()
torch.cuda.synchronize()
l0 = _lib.launch_count()
step()
torch.cuda.synchronize()
print("aesr launches per step:", _lib.launch_count() - l0)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
print("eager: host enqueue %.3f ms/step, device %.3f ms/step" % (1e3 * t_enq / steps, e0.elapsed_time(e1) / steps))

try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        res = step()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("graph replay: device %.3f ms/step (Adam step count frozen in the capture -- probe only)" %
          (e0.elapsed_time(e1) / steps))
except Exception as ex:      # noqa: BLE001
    print("graph capture failed:", repr(ex)[:600])
