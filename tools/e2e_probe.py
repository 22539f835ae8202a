"""Where does the host-buffer path (synthesis.HostPipeline) spend its step?  Times its parts in isolation:
PCIe copies (contiguous / strided rows, alone and both directions at once), the host worker's clamp, the enqueue
cost of one synthesize_volumes call on the host, and the pipeline with parts switched off.

  python tools/e2e_probe.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aesr_oracle as O  # noqa: E402
from superresolution_aniso_mri_b200 import ops, synthesis  # noqa: E402
from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI  # noqa: E402

dev = torch.device("cuda:0")
V, Z, S, NI = 64, 10, 128, 6
Zo = (Z - 1) * (NI + 1) + 1
args = O.default_args(S, 32)
margs = dict(args)
margs["device"] = "cuda:0"
model = VanillaACAI(margs)
model.load_state_dict(O.calibrated_state(args))
model.eval()
ar = O.alpha_range_for(NI)
host_in = torch.rand(V, Z, S, S).pin_memory()
host_out = torch.empty(V, Zo, S, S).pin_memory()
d_in = host_in.to(dev)
d_out = torch.empty(V, Zo, S, S, device=dev)


def wall(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


sl, vol = S * S * 4, Zo * S * S * 4
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def d2h_full():
    host_out.copy_(d_out, non_blocking=True)


def d2h_rows():
    ops.copy_rows_async(host_out, d_out, sl, sl, outer=V, dst_outer_stride=vol, src_outer_stride=vol, rows=Z - 1,
                        dpitch=(NI + 1) * sl, spitch=(NI + 1) * sl, width=NI * sl, stream=torch.cuda.current_stream())


def h2d():
    d_in.copy_(host_in, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d2h_rows()
    with torch.cuda.stream(s2):
        h2d()


t = wall(d2h_full)
print("D2H contiguous %5.1f MB: %.3f ms  %.1f GB/s" % (host_out.numel() * 4 / 1e6, t, host_out.numel() * 4 / t / 1e6))
t = wall(d2h_rows)
nb = V * (Z - 1) * NI * sl
print("D2H synthesized rows %5.1f MB: %.3f ms  %.1f GB/s" % (nb / 1e6, t, nb / t / 1e6))
t = wall(h2d)
print("H2D %5.1f MB: %.3f ms  %.1f GB/s" % (host_in.numel() * 4 / 1e6, t, host_in.numel() * 4 / t / 1e6))
t = wall(both)
print("D2H rows + H2D concurrently: %.3f ms" % t)
t0 = time.perf_counter()
for _ in range(10):
    torch.clamp(host_in, 0.0, 1.0, out=host_out[:, ::NI + 1])
print("host clamp of the kept slices (%.1f MB): %.3f ms" % (host_in.numel() * 4 / 1e6, (time.perf_counter() - t0) * 100))
# enqueue cost: host time of one synthesize_volumes call while the GPU is kept busy (no sync inside the loop)
for _ in range(3):
    synthesis.synthesize_volumes(model, d_in, ar, out=d_out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    synthesis.synthesize_volumes(model, d_in, ar, out=d_out)
t_host = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 10 * 1e3
print("synthesize_volumes, 64 volumes: host enqueue %.3f ms per call, device %.3f ms per call" % (t_host, t_all))
for groups in (1, 2, 4):
    for kept in (True, False):
        pipe = synthesis.HostPipeline(model, V, Z, S, S, ar, groups=groups, host_kept=kept)
        for _ in range(3):
            pipe.run(host_in, host_out, wait=False)
        pipe.synchronize()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            pipe.run(host_in, host_out, wait=False)
        t_enq = (time.perf_counter() - t0) / 10 * 1e3
        pipe.synchronize()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 10 * 1e3
        print("HostPipeline groups=%d host_kept=%s: %.3f ms per step (host enqueue %.3f ms)" % (groups, kept, t, t_enq))

# ---- independent streams: does compute slow down the copies (or the reverse) when nothing orders them?
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
reps = 10
with torch.cuda.stream(sa):
    evs[0].record()
    for _ in range(reps):
        synthesis.synthesize_volumes(model, d_in, ar, out=d_out)
    evs[1].record()
with torch.cuda.stream(sb):
    evs[2].record()
    for _ in range(reps):
        ops.copy_rows_async(host_out, d_out, sl, sl, outer=V, dst_outer_stride=vol, src_outer_stride=vol, rows=Z - 1,
                            dpitch=(NI + 1) * sl, spitch=(NI + 1) * sl, width=NI * sl, stream=sb)
        d_in.copy_(host_in, non_blocking=True)
    evs[3].record()
torch.cuda.synchronize()
print("unordered, concurrent: compute %.3f ms per step, D2H rows + H2D %.3f ms per step" % (
    evs[0].elapsed_time(evs[1]) / reps, evs[2].elapsed_time(evs[3]) / reps))

# ---- timeline of the pipeline in steady state (events on the compute and the download streams)
pipe = synthesis.HostPipeline(model, V, Z, S, S, ar, groups=2)
for _ in range(3):
    pipe.run(host_in, host_out, wait=False)
pipe.synchronize()
torch.cuda.synchronize()
pipe.trace = []
t_ref = torch.cuda.Event(enable_timing=True)
t_ref.record()
for _ in range(4):
    pipe.run(host_in, host_out, wait=False)
pipe.synchronize()
torch.cuda.synchronize()
for tag, grp, ev in pipe.trace:
    print("  %-14s group@%-3d %8.3f ms" % (tag, grp, t_ref.elapsed_time(ev)))
