#!/bin/bash
# r02q: weight-gradient time vs number of CTAs (each CTA ends with Cout x Cin x taps fp32 atomics)
for c in 148 96 64 36 18; do
AESR_WGRAD_CTAS=$c timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02q_$c.txt 2>&1
echo "== AESR_WGRAD_CTAS=$c"; grep wgrad gpurun_out/train_layers_r02q_$c.txt | awk '{printf "%s ", $3}'; echo; tail -1 gpurun_out/train_layers_r02q_$c.txt | grep -o "'wgrad3x3': [0-9.]*"
done
