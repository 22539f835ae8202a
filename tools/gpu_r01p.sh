#!/bin/bash
timeout 500 python -m pytest tests -m gpu -q -x -k "dhcp or sweep_downsample or brain_config or find_best or vif or compute_metrics or lpips_metric" > gpurun_out/pytest_r01p.log 2>&1; tail -5 gpurun_out/pytest_r01p.log
timeout 200 python tools/gpu_eager_baseline.py > gpurun_out/gpu_eager_r01p.log 2>&1; tail -3 gpurun_out/gpu_eager_r01p.log
