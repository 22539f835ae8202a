#!/bin/bash
# r02b: paired stores only for BN >= 64, stem with L2 prefetch -- parity, layer A/B, stem timing
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_r02b.log
PREV=$PWD/superresolution_aniso_mri_b200/lib/libaesr_b200_prev.so
AESR_B200_LIB=$PREV timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r02b_prev.txt 2>&1
timeout 300 python tools/layer_times.py --enc 640 --dec 3456 --reps 10 > gpurun_out/layers_r02b_new.txt 2>&1
paste gpurun_out/layers_r02b_prev.txt gpurun_out/layers_r02b_new.txt | cut -c1-75,124-160
timeout 300 python tools/head_sweep.py --stem-only 2>&1 | grep -v Warn | tee gpurun_out/stem_sweep_r02b.txt
