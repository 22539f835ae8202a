#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "decoder_tail" > gpurun_out/pytest_r01i.log 2>&1; tail -12 gpurun_out/pytest_r01i.log
timeout 240 python tools/conv_sweep.py > gpurun_out/conv_sweep_r01i.log 2>&1; head -12 gpurun_out/conv_sweep_r01i.log
