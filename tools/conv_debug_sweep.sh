#!/bin/bash
# Stage-isolation sweep of the halo conv kernel (AESR_CONV_DEBUG bits, see conv3x3_tc.cuh): which stage bounds a tile?
#   2 no activation TMA | 4 no stores | 8 no TMEM loads | 32 no MMAs | 64 no epilogue math
for shape in ${SHAPES:-"32 32 128 252 0" "64 64 65 256 1"}; do
  for dbg in ${DBGS:-0 2 32 34 76 78 108 110}; do
    echo -n "debug=$dbg  "
    AESR_CONV_DEBUG=$dbg python tools/bench_conv.py $shape 1 10
  done
done
