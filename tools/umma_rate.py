"""tcgen05.mma cost vs operand layout and accumulator rotation (diagnostic): cycles per MMA (M=128, K=16, fp16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

lib = _lib.lib_for_device(0)
out = torch.zeros(1, dtype=torch.int64, device="cuda:0")
iters = 2000
print("kc  N  nacc pitch shift adv : cycles/MMA")
for kc in (32, 64):
    for N in (32, 64, 128, 256):
        for nacc in (1, 2, 4, 8):
            if nacc * N > 512:
                continue
            for pitch, shift, adv in ((8, 0, 0), (10, 1, 1)):
                _lib.check(lib.aesr_probe_umma_rate(out.data_ptr(), N, kc, pitch, shift, iters, adv, nacc,
                                                    torch.cuda.current_stream().cuda_stream), "probe")
                torch.cuda.synchronize()
                print("%2d %3d  %2d   %3d   %3d  %3d : %.1f" % (kc, N, nacc, pitch, shift, adv, out.item() / iters))
