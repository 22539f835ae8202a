#!/bin/bash
# r02u: ncu launch list of the training step (true kernel durations, no launch gaps) + final default bench
timeout 300 python tools/train_layer_times.py --reps 1 > gpurun_out/plain_train_r02u.log 2>&1; echo "plain rc $?"
AESR_TRAIN_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_r02u.csv python tools/train_layer_times.py --reps 1 > gpurun_out/ncu_train_r02u.log 2>&1; echo "ncu rc $?"
timeout 600 python bench.py > gpurun_out/bench_r02u.json 2> gpurun_out/bench_r02u.err; echo "bench rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02u.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_r02u.log
