"""tcgen05.mma cost vs operand layout, accumulator rotation, number of concurrently issuing SMs and operand data
(diagnostic): SM cycles per MMA (M=128, K=16, fp16) and wall-clock ns per MMA (=> effective SM clock).

  python tools/umma_rate.py [--full]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

lib = _lib.load_probe()          # diagnostic build (include/aesr_b200_probe.h)
_lib.check(lib.aesr_init(0), "init")
out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda:0")
iters = 20000
full = "--full" in sys.argv


def run(kc, N, nacc, pitch, shift, adv, grid, rnd):
    out.zero_()
    for _ in range(2):
        _lib.check(lib.aesr_probe_umma_rate(out.data_ptr(), N, kc, pitch, shift, iters, adv, nacc, grid, rnd,
                                            torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    r = out.view(-1, 2)[:grid].double()
    cyc, ns = r[:, 0].max().item() / iters, r[:, 1].max().item() / iters
    print("%2d %3d  %2d   %3d   %3d  %3d  %4d  %s : %6.1f cycles/MMA  %6.1f ns/MMA  (%4.0f MHz effective)"
          % (kc, N, nacc, pitch, shift, adv, grid, "rand" if rnd else "zero", cyc, ns, 1e3 * cyc / ns))


if "--pattern" in sys.argv:
    # the conv kernel's issue loop in isolation: cycles per MMA for the layer shapes of the ACDC pipeline
    print("BN kc  T variant lag grid data : cycles/MMA (probe rate: N=32 40.1, 64 48.0, 128 64.0)")
    its = 400
    for BN, kc, T in ((32, 32, 4), (64, 32, 2), (64, 64, 2), (128, 32, 1), (128, 64, 1), (64, 64, 1)):
        per = T * 9 * (kc // 16)
        for variant, lag in ((0, 1), (1, 1), (3, 3), (7, 3), (8, 1), (16, 1), (32, 1), (56, 1)):
            for grid, rnd in ((1, 0),):
                out.zero_()
                for _ in range(2):
                    _lib.check(lib.aesr_probe_umma_pattern(out.data_ptr(), BN, kc, T, its, variant, lag, grid, rnd,
                                                           torch.cuda.current_stream().cuda_stream), "pattern")
                torch.cuda.synchronize()
                r = out.view(-1, 2)[:grid].double()
                cyc, ns = r[:, 0].max().item() / (its * per), r[:, 1].max().item() / (its * per)
                print("%3d %2d  %d    %d     %d  %4d  %s : %6.1f cycles/MMA  %6.1f ns/MMA" %
                      (BN, kc, T, variant, lag, grid, "rand" if rnd else "zero", cyc, ns))
    sys.exit(0)
print("kc  N  nacc pitch shift adv  grid data")
for kc in (32, 64):
    for N in (32, 64, 128, 256):
        for grid in (1, 148):
            for rnd in (0, 1):
                run(kc, N, min(4, 512 // N), 10, 1, 1, grid, rnd)
        if full:
            for nacc in (1, 2, 4, 8):
                if nacc * N <= 512:
                    run(kc, N, nacc, 8, 0, 0, 1, 0)
