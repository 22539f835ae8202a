"""What would programmatic dependent launch buy the graph-replayed training step?  Device time per kernel of a captured chain
of dependent launches, with and without the PDL attribute (diagnostic library, include/aesr_b200_probe.h).

  python tools/launch_gap_probe.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import _lib  # noqa: E402

lib = _lib.load_probe()
_lib.check(lib.aesr_init(0), "init")
out = ctypes.c_float(0.0)
print("kernels ctas spin_cycles :  us/kernel plain   us/kernel PDL   (spin at ~1.9 GHz: 0 / 9500 / 38000 cycles = 0 / 5 / 20 us)")
for ctas in (1, 148, 592):
    for spin in (0, 9500, 38000):
        row = []
        for pdl in (0, 1):
            _lib.check(lib.aesr_probe_launch_gap(ctypes.byref(out), 128, ctas, spin, pdl, 20), "probe")
            row.append(out.value)
        print("%7d %4d %11d :  %14.2f  %14.2f   gap %.2f -> %.2f us" % (128, ctas, spin, row[0], row[1],
              row[0] - spin / 1900.0, row[1] - spin / 1900.0))
