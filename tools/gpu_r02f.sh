#!/bin/bash
# r02f: scalar constant-operand FFMA head epilogue vs packed fp32x2 (previous build) -- parity + stage isolation A/B
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_head or golden or synth" > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_r02f.log
PREV=$PWD/superresolution_aniso_mri_b200/lib/libaesr_b200_prev.so
echo "--- packed (previous build)"; AESR_B200_LIB=$PREV timeout 300 python tools/head_sweep.py 2>&1 | grep -v Warn | head -14 | tee gpurun_out/head_sweep_r02f_packed.txt
echo "--- scalar"; timeout 300 python tools/head_sweep.py 2>&1 | grep -v Warn | head -14 | tee gpurun_out/head_sweep_r02f_scalar.txt
