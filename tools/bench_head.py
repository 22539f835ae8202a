"""Single-launch timing of the fused decoder tail: python tools/bench_head.py N [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import ops  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
dt = torch.float16
x = torch.randn(n, 64, 64, 32, device=dev).to(dt)
wp = ops.pack_conv3x3_weight_up2fold(torch.randn(32, 32, 3, 3, device=dev) * 0.05, dtype=dt)
b = torch.zeros(32, device=dev)
hw = torch.randn(9, 32) * 0.1
out = torch.empty(n, 64, 64, 16, device=dev)
for _ in range(2):
    ops.conv3x3_up2_head(x, wp, b, hw, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv3x3_up2_head(x, wp, b, hw, out=out)
e1.record()
torch.cuda.synchronize()
print("up2+head n=%d: %.3f ms" % (n, e0.elapsed_time(e1) / reps))
# reference points for the memory system: pure write, pure read+write
big = torch.empty(256 << 20, dtype=torch.float32, device=dev)
for name, fn, nbytes in (("fill 1 GiB", lambda: big.fill_(1.0), big.numel() * 4),
                         ("copy 0.5 GiB", lambda: big[:128 << 20].copy_(big[128 << 20:]), big.numel() * 4)):
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.1f GB/s" % (name, nbytes * 5 / e0.elapsed_time(e1) / 1e6))
