#!/bin/bash
# Round profile set (run under gpurun): plain run, ncu launch lists (inference step, training step), ncu --set full of the
# conv kernels of one inference step.  Reports stay under /tmp on the box (they can exceed gpurun's 64 MiB return limit);
# only the text / json summaries come back in gpurun_out/.
#   bash tools/profile_round.sh TAG
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file /tmp/launches_${TAG}.csv $CMD > /tmp/ncu_launches_${TAG}.log 2>&1
echo "launch list rc $?"
python tools/ncu_summary.py launches /tmp/launches_${TAG}.csv gpurun_out/${TAG}_launches.md
# training step: launch list of ONE eager step (tools/train_ncu.py brackets it with the profiler range)
python tools/train_ncu.py > /tmp/train_plain_${TAG}.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_train_${TAG}.csv python tools/train_ncu.py > /tmp/ncu_train_${TAG}.log 2>&1
echo "train launch list rc $?"
python tools/ncu_summary.py launches /tmp/launches_train_${TAG}.csv gpurun_out/${TAG}_train_launches.md
# full capture of the 10 conv launches of one inference step (5 encoder layers, dec.0 on the latents, 4 decoder launches); the
# warm-up steps are skipped (3 warm-up + 3 e2e warm-up... : profile the LAST step's convs: skip 30, take 10)
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 30 -c 10 -f -o /tmp/prof_conv_${TAG} $CMD > /tmp/ncu_full_${TAG}.log 2>&1
echo "full capture rc $?"
python tools/ncu_summary.py full /tmp/prof_conv_${TAG}.ncu-rep gpurun_out/${TAG}_conv_full.md
ls -la /tmp/prof_conv_${TAG}.ncu-rep gpurun_out/${TAG}_*
