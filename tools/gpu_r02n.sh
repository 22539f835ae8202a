#!/bin/bash
# r02n: SM-driven download probe; collapsed VGG conv1_1; training parity + LPIPS tests + bench
# (the SM-driven download probe that ran here, tools/e2e_d2h_probe.py, was removed with the experiment; its output is kept
#  as profiles/r02n_e2e_sm_download_probe.txt)
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_eval.py -m gpu -x -q > gpurun_out/pytest_r02n.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02n.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02n.txt 2>&1; tail -1 gpurun_out/train_layers_r02n.txt | cut -c1-420
