#!/bin/bash
# r02a: sector-complete paired stores in the lean epilogue -- parity + per-layer A/B against the previous build + stem ncu
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02a.log
PREV=$PWD/superresolution_aniso_mri_b200/lib/libaesr_b200_prev.so
for i in 1 2; do
AESR_B200_LIB=$PREV timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r02a_prev$i.txt 2>&1
timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r02a_new$i.txt 2>&1
done
paste gpurun_out/layers_r02a_prev2.txt gpurun_out/layers_r02a_new2.txt | cut -c1-75,124-160
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 3 -c 1 -o gpurun_out/prof_stem_r02a python tools/head_sweep.py --stem-only > gpurun_out/ncu_stem_r02a.log 2>&1; echo "ncu rc $?"
