#!/bin/bash
# r01s: channel-chunk-major head epilogue (dec.12 + head) -- parity, A/B layer times against the previous build, bench
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r01s.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r01s.log
PREV=$PWD/superresolution_aniso_mri_b200/lib/libaesr_b200_prev.so
for i in 1 2; do
AESR_B200_LIB=$PREV timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r01s_prev$i.txt 2>&1
timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r01s_new$i.txt 2>&1
done
tail -8 gpurun_out/layers_r01s_prev2.txt; tail -8 gpurun_out/layers_r01s_new2.txt
timeout 400 python bench.py --steps 10 --no-train --cpu-sample 1 > gpurun_out/bench_r01s.json 2> gpurun_out/bench_r01s.err; echo "bench rc $?"
