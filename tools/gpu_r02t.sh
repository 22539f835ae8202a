#!/bin/bash
# r02t: BN statistics accumulated per CTA in shared memory -- full GPU suite, per-launch training times, bench
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02t.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02t.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02t.txt 2>&1; sed -n 1,32p gpurun_out/train_layers_r02t.txt | grep "conv3x3" | cut -c1-100; tail -1 gpurun_out/train_layers_r02t.txt | cut -c1-300
timeout 600 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r02t.json 2> gpurun_out/bench_r02t.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02t.json"))
t = d["train"]
print("infer", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
