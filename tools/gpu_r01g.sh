#!/bin/bash
timeout 240 python tools/conv_sweep.py > gpurun_out/conv_sweep_r01g.log 2>&1; tail -3 gpurun_out/conv_sweep_r01g.log
timeout 300 python bench.py --steps 20 --no-train > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; tail -c 300 gpurun_out/bench_r01g.err
