#!/bin/bash
# r01r: full GPU suite + bench with the shared-memory stem filter
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01r.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r01r.log
timeout 400 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r01r.json 2> gpurun_out/bench_r01r.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r01r.err
