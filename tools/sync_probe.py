"""mbarrier round-trip latency between two warps (diagnostic; see probe.cuh sync_probe_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

lib = _lib.load_probe()          # diagnostic build (include/aesr_b200_probe.h)
_lib.check(lib.aesr_init(0), "init")
out = torch.zeros(1, dtype=torch.int64, device="cuda:0")
iters = 2000
for mode in range(8):
    _lib.check(lib.aesr_probe_sync(out.data_ptr(), iters, mode, torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    print("mode %d (%s signal, %s, %d waiting warps): %.1f cycles per round trip" % (
        mode, "tcgen05.commit" if mode & 1 else "mbarrier.arrive", "test_wait poll" if mode & 2 else "try_wait",
        3 if mode & 4 else 1, out.item() / iters))
