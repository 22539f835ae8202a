#!/bin/bash
# r01u: stage isolation of dec.12 + head for both head variants; warp-MMA stem parity + A/B
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r01u.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r01u.log
timeout 600 python tools/head_sweep.py > gpurun_out/head_sweep_r01u.txt 2>&1; echo "sweep rc $?"; cat gpurun_out/head_sweep_r01u.txt
