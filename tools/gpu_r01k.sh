#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01k.log 2>&1; tail -3 gpurun_out/pytest_r01k.log
timeout 400 python bench.py > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err; tail -c 300 gpurun_out/bench_r01k.err
timeout 200 python bench.py --no-train --steps 10 --groups 1 --cpu-sample 1 > gpurun_out/bench_r01k_groups1.json 2>&1
timeout 600 bash tools/profile_round.sh r01k > gpurun_out/profile_round_r01k.log 2>&1; tail -5 gpurun_out/profile_round_r01k.log
