#!/bin/bash
# r01x: synthesized-slices-only download in HostPipeline (host writes the kept slices) -- parity + e2e bench, pure-write bandwidth probe
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline or stem" > gpurun_out/pytest_r01x.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r01x.log
timeout 400 python bench.py --steps 20 --no-train --cpu-sample 1 > gpurun_out/bench_r01x.json 2> gpurun_out/bench_r01x.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r01x.err
timeout 400 python bench.py --steps 20 --no-train --cpu-sample 1 --groups 4 > gpurun_out/bench_r01x_g4.json 2> gpurun_out/bench_r01x_g4.err; echo "bench rc $?"
timeout 120 python - > gpurun_out/write_bw_r01x.txt 2>&1 <<'PY'
import torch
dev = torch.device("cuda:0")
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mb in (692, 2048, 4096):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
    t = timed(lambda: a.zero_())
    print("fill  %5d MB: %.3f ms  %.0f GB/s (write only)" % (mb, t, mb * 1.048576 / t))
    t = timed(lambda: b.copy_(a))
    print("copy  %5d MB: %.3f ms  %.0f GB/s (read + write)" % (mb, t, 2 * mb * 1.048576 / t))
    t = timed(lambda: a.sum())
    print("sum   %5d MB: %.3f ms  %.0f GB/s (read only)" % (mb, t, mb * 1.048576 / t))
    del a, b
PY
cat gpurun_out/write_bw_r01x.txt
