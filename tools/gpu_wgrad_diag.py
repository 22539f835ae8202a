"""wgrad bring-up: tcgen05 (algo 1) vs CUDA-core (algo 2) vs torch fp32 on the rounded operands; timing."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import build, ops_train as T  # noqa: E402

build.build_library()
dev = torch.device("cuda:0")


def ref_wgrad(g, x):
    """dW[co][ci][dy][dx] = sum g[n,y,x,co] * xpad[n,y+dy,x+dx,ci]"""
    gf = g.float().permute(0, 3, 1, 2).cpu()
    xf = x.float().permute(0, 3, 1, 2).cpu()
    xs = xf.requires_grad_(False)
    w = torch.zeros(gf.shape[1], xf.shape[1], 3, 3, requires_grad=True)
    y = F.conv2d(xs, w, padding=1)
    (y * gf).sum().backward()
    return w.grad, gf.sum(dim=(0, 2, 3))


for (cin, cout, n, h, w) in ((64, 64, 1, 16, 8), (64, 64, 2, 32, 32), (32, 32, 2, 40, 24), (32, 32, 3, 130, 130),
                             (64, 32, 2, 65, 65), (32, 64, 2, 64, 64), (128, 64, 2, 32, 32), (64, 128, 2, 32, 32),
                             (128, 128, 2, 32, 32)):
    for xdt in ((torch.float16,) if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else (torch.bfloat16,)):
        try:
            gen = torch.Generator().manual_seed(cin + cout + h)
            g = (torch.randn(n, h, w, cout, generator=gen) * 0.1).to(torch.bfloat16).to(dev)
            x = torch.randn(n, h, w, cin, generator=gen).to(xdt).to(dev)
            want, wb = ref_wgrad(g, x)
            res = {}
            for algo in (1, 2):
                dW = torch.zeros(cout, cin, 3, 3, device=dev)
                db = torch.zeros(cout, device=dev)
                T.wgrad3x3(g, x, dW, db, algo=algo)
                torch.cuda.synchronize()
                res[algo] = ((dW.cpu() - want).abs().max().item(), (db.cpu() - wb).abs().max().item())
            print("wgrad %3d->%3d n%d %3dx%3d x=%s: tc err %.4f (db %.4f) | cuda err %.4f | ref max %.2f %s"
                  % (cin, cout, n, h, w, str(xdt)[6:], res[1][0], res[1][1], res[2][0], want.abs().max().item(),
                     "OK" if res[1][0] < 2e-3 * max(1.0, want.abs().max().item()) + 1e-2 else "WRONG"))
        except Exception:
            traceback.print_exc()

# timing at ACDC step shapes
for name, cin, cout, hw, n in (("E1", 32, 32, 130, 24), ("E4", 64, 64, 65, 24), ("E6", 128, 128, 32, 24),
                               ("D0", 128, 64, 32, 24), ("D2", 64, 32, 64, 24), ("D4", 32, 32, 128, 24)):
    g = (torch.randn(n, hw, hw, cout, device=dev) * 0.1).to(torch.bfloat16)
    x = torch.randn(n, hw, hw, cin, device=dev).to(torch.bfloat16)
    dW = torch.zeros(cout, cin, 3, 3, device=dev)
    db = torch.zeros(cout, device=dev)
    out = []
    for algo in (1, 2):
        for _ in range(2):
            T.wgrad3x3(g, x, dW, db, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            T.wgrad3x3(g, x, dW, db, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 5)
    print("%s %d->%d @%d n=%d: tc %.3f ms | cuda %.3f ms" % (name, cin, cout, hw, n, out[0], out[1]))
