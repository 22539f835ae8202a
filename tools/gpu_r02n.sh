#!/bin/bash
# r02n: SM-driven download probe; collapsed VGG conv1_1; training parity + LPIPS tests + bench
timeout 300 python tools/e2e_d2h_probe.py 2>&1 | grep -v Warn > gpurun_out/e2e_d2h_probe_r02n.txt; cat gpurun_out/e2e_d2h_probe_r02n.txt
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_eval.py -m gpu -x -q > gpurun_out/pytest_r02n.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02n.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02n.txt 2>&1; tail -1 gpurun_out/train_layers_r02n.txt | cut -c1-420
