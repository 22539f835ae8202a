"""Single-layer conv timing for profiling: python tools/bench_conv.py CIN COUT HW N MODE ALGO [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

cin, cout, hw, n, mode, algo = [int(a) for a in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = torch.device("cuda:0")
dt = torch.float16
x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, dtype=dt)
b = torch.zeros(cout, device=dev)
out = torch.empty(ops.conv_out_shape(n, hw, hw, cout, mode) if mode != 5 else (n, 2 * hw, 2 * hw, cout // 4), dtype=torch.float32 if mode == 3 else dt, device=dev)
for _ in range(2):
    ops.conv3x3(x, wp, b, act=1, out_mode=mode, out=out, algo=algo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv3x3(x, wp, b, act=1, out_mode=mode, out=out, algo=algo)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("conv %d->%d @%d n=%d mode=%d algo=%d: %.3f ms, %.1f TFLOP/s" % (cin, cout, hw, n, mode, algo, ms, 2.0 * n * hw * hw * 9 * cin * cout / ms / 1e9))
