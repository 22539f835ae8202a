#!/bin/bash
# round-1d diagnostics: fresh per-layer times, chunk-size sweep, stage isolation of the decoder-shaped layers
python tools/layer_times.py > gpurun_out/layers_r01d.log 2>&1
for c in 256 512 1024 3456; do
  python bench.py --no-train --steps 10 --chunk $c --cpu-sample 1 > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err
done
{
while read -r shape; do
  for dbg in 0 32 76 110; do
    echo -n "debug=$dbg  "
    AESR_CONV_DEBUG=$dbg python tools/bench_conv.py $shape 1 10
  done
done <<'SH'
128 64 32 252 0
64 64 32 252 0
64 128 32 252 5
32 32 64 252 0
32 32 130 256 1
64 64 65 256 1
SH
} > gpurun_out/debug_sweep_r01d.log 2>&1
