#!/bin/bash
for c in 576 864 1152 1728 4096; do
  timeout 200 python bench.py --no-train --steps 10 --chunk $c --cpu-sample 1 > gpurun_out/bench_r01j_chunk$c.json 2> gpurun_out/bench_r01j_chunk$c.err
done
AESR_HEAD_TC=0 timeout 200 python bench.py --no-train --steps 10 --cpu-sample 1 > gpurun_out/bench_r01j_headcc.json 2>&1
for g in 2 8; do
  timeout 200 python bench.py --no-train --steps 10 --groups $g --cpu-sample 1 > gpurun_out/bench_r01j_groups$g.json 2>&1
done
