#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-train --cpu-sample 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"stem_conv|head_gather|lerp_pairs_act" -s 9 -c 3 -o gpurun_out/prof_mem_r01q $CMD > gpurun_out/ncu_mem_r01q.log 2>&1
echo "ncu rc $?"
timeout 300 python bench.py --no-train --steps 10 --cpu-sample 1 > gpurun_out/bench_r01q.json 2> gpurun_out/bench_r01q.err; tail -c 200 gpurun_out/bench_r01q.err
