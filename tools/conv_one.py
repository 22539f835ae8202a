"""One 3x3 conv layer at a given launch size, a few launches (ncu target).  python tools/conv_one.py cin cout hw n mode [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

cin, cout, hw, n, mode = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device("cuda:0")
dt = torch.float16
x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
cb = cout // 4 if mode == 5 else cout
b = torch.zeros(cb, device=dev)
sc, sh = torch.ones(cb, device=dev), torch.zeros(cb, device=dev)
wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, dtype=dt)
out = torch.empty(ops.conv_out_shape(n, hw, hw, cout, mode), dtype=dt, device=dev)
for _ in range(reps):
    ops.conv3x3(x, wp, b, act=1, scale=sc, shift=sh, out_mode=mode, out=out)
torch.cuda.synchronize()
print("ok")
