#!/bin/bash
# r01t: decoder head on warp-level mma.sync fragments -- parity of the variant, whole-pipeline parity with it on, A/B timing
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_head" > gpurun_out/pytest_r01t_head.log 2>&1; echo "pytest head rc $?"; tail -3 gpurun_out/pytest_r01t_head.log
AESR_HEAD_MMA=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01t_mma.log 2>&1; echo "pytest mma rc $?"; tail -3 gpurun_out/pytest_r01t_mma.log
for i in 1 2; do
timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r01t_cuda$i.txt 2>&1
AESR_HEAD_MMA=1 timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r01t_mma$i.txt 2>&1
done
tail -4 gpurun_out/layers_r01t_cuda2.txt; tail -4 gpurun_out/layers_r01t_mma2.txt
AESR_HEAD_MMA=1 timeout 400 python bench.py --steps 10 --no-train --cpu-sample 1 > gpurun_out/bench_r01t_mma.json 2> gpurun_out/bench_r01t_mma.err; echo "bench rc $?"
timeout 400 python bench.py --steps 10 --no-train --cpu-sample 1 > gpurun_out/bench_r01t_cuda.json 2> gpurun_out/bench_r01t_cuda.err; echo "bench rc $?"
