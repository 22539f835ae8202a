"""Race hunt: every conv configuration is deterministic (no atomics unless stats are requested), so repeated launches on
the same inputs must be bit-identical.  Prints the number of mismatching repeats per configuration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_aniso_mri_b200 import build, ops  # noqa: E402

build.build_library()
dev = torch.device("cuda:0")
cases = [  # cin, cout, hw, n, act, mode
    (32, 32, 130, 24, 1, 0), (32, 32, 130, 24, 1, 1), (32, 64, 65, 24, 1, 0), (64, 64, 65, 24, 1, 1),
    (64, 128, 32, 24, 1, 0), (128, 128, 32, 24, 0, 3), (128, 64, 32, 12, 1, 0), (64, 64, 32, 12, 1, 2),
    (64, 32, 64, 12, 1, 0), (32, 32, 64, 12, 1, 2), (32, 32, 128, 12, 1, 0),
    (64, 64, 128, 24, 2, 4), (64, 128, 64, 24, 2, 0), (128, 128, 64, 24, 2, 4), (128, 256, 32, 24, 2, 0),
    (256, 256, 32, 24, 2, 0), (256, 256, 32, 24, 2, 4), (256, 512, 16, 24, 2, 0), (512, 512, 16, 24, 2, 4),
    (512, 512, 8, 24, 2, 0),
]
for dt in (torch.float16, torch.bfloat16):
    for cin, cout, hw, n, act, mode in cases:
        x = torch.randn(n, hw, hw, cin, device=dev).to(dt)
        wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, dtype=dt)
        b = torch.randn(cout, device=dev) * 0.1
        ref = None
        bad = 0
        for rep in range(12):
            out = ops.conv3x3(x, wp, b, act=act, out_mode=mode)
            outs = out if isinstance(out, tuple) else (out,)
            torch.cuda.synchronize()
            if ref is None:
                ref = [o.clone() for o in outs]
            else:
                if not all(torch.equal(o, r) for o, r in zip(outs, ref)):
                    bad += 1
        print("%s %3d->%3d @%3d n%d act%d mode%d : %d/11 repeats differ %s"
              % (str(dt)[6:], cin, cout, hw, n, act, mode, bad, "" if bad == 0 else "<<<<<< RACE"))
