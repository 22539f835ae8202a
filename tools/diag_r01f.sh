#!/bin/bash
# stage isolation at LARGE n (fixed launch cost amortised): AESR_CONV_DEBUG bits 2 no TMA loads | 4 no stores |
# 8 no TMEM loads | 32 no MMAs | 64 no epilogue math (and no stores)
{
while read -r shape; do
  for dbg in 0 4 8 12 64 76 32 2 34 110; do
    echo -n "debug=$dbg  "
    AESR_CONV_DEBUG=$dbg python tools/bench_conv.py $shape 1 5
  done
done <<'SH'
64 128 32 3456 5
32 32 64 3456 0
64 64 32 3456 0
32 64 65 640 0
32 32 130 640 1
128 128 32 640 0
SH
for dbg in 0 32 2 34; do echo -n "debug=$dbg  "; AESR_CONV_DEBUG=$dbg python tools/bench_head.py 3456 5 | head -1; done
} > gpurun_out/debug_sweep_r01f.log 2>&1
