#!/bin/bash
# halo (resident filter bank) vs streamed kernel on the LPIPS-VGG trunk shapes of the ACDC training step (n = 24 forward, 12 backward)
for shape in "64 64 128 24" "64 128 64 24" "128 128 64 24" "128 256 32 24" "256 256 32 24" "256 512 16 24" "512 512 16 24" \
             "64 64 128 12" "128 64 64 12" "128 128 64 12" "256 128 32 12" "256 256 32 12" "512 256 16 12" "512 512 16 12" "512 512 8 12" "512 512 8 24"; do
  set -- $shape
  for algo in 1 2; do
    python tools/bench_conv.py $1 $2 $3 $4 0 $algo 20 2>&1 | tail -1
  done
done
