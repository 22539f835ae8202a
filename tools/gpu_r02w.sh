#!/bin/bash
# r02w: e0 backward with 16-byte loads and one set of atomics per block -- training parity + bench
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_r02w.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02w.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02w.txt 2>&1; tail -1 gpurun_out/train_layers_r02w.txt | cut -c1-330
timeout 600 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02w.json"))
t = d["train"]
print("infer", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
