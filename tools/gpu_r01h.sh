#!/bin/bash
timeout 120 python tools/umma_rate.py > gpurun_out/umma_rate_r01h.log 2>&1; tail -4 gpurun_out/umma_rate_r01h.log
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "stem or fold or head or golden or spec or pipeline" > gpurun_out/pytest_r01h.log 2>&1; tail -2 gpurun_out/pytest_r01h.log
timeout 300 python bench.py --steps 20 --no-train > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err; tail -c 300 gpurun_out/bench_r01h.err
