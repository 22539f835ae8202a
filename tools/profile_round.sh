#!/bin/bash
# Round profile set (run under gpurun): plain run, ncu launch list, ncu --set full of the conv kernels of one small step.
#   bash tools/profile_round.sh TAG
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-train --cpu-sample 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc $?"
# full capture of the 10 conv launches of one step (5 encoder layers, dec.0 on the latents, 4 decoder launches); the
# 3 warm-up steps (30 conv launches) are skipped
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 30 -c 10 -o gpurun_out/prof_conv_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc $?"
ls -la gpurun_out/*${TAG}*
