"""TMEM read rate and shuffle rate per SM (diagnostic): bytes per clock of tcgen05.ld.32x32b.x16 with 4 / 8 / 16 warps per CTA
(one / two / four warps per TMEM lane quarter), and cycles per fp32 shuffle.

  python tools/tmem_ld_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

lib = _lib.load_probe()
_lib.check(lib.aesr_init(0), "init")
out = torch.zeros(148, dtype=torch.int64, device="cuda:0")
sink = torch.zeros(512, device="cuda:0")
iters = 4000
for grid in (1, 148):
    for mode, what in ((0, "tcgen05.ld x16"), (1, "shfl"), (2, "both")):
        for nw in (4, 8, 16):
            for _ in range(2):
                _lib.check(lib.aesr_probe_tmem_ld(out.data_ptr(), nw, iters, mode, grid, sink.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream), "probe")
            torch.cuda.synchronize()
            cyc = out[:grid].max().item()
            line = "grid %3d  %-14s %2d warps: %9d cycles" % (grid, what, nw, cyc)
            if mode != 1:
                line += "  %6.1f B/clk/SM TMEM read" % (nw * iters * 4 * 2048 / cyc)
            if mode != 0:
                line += "  %5.2f clk per warp-shuffle (SM-wide)" % (cyc / (nw * iters * 64))
            print(line)
