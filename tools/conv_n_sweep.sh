#!/bin/bash
# fixed vs per-tile cost of the halo conv kernel: t(n) = a + b * tiles
for dbg in 0 110; do
  for n in 1 8 32 64 128 252 504; do
    echo -n "debug=$dbg  "
    AESR_CONV_DEBUG=$dbg python tools/bench_conv.py 32 32 128 $n 0 1 20
  done
done
