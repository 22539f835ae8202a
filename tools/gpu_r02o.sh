#!/bin/bash
# r02o: wgrad with the dx taps folded into N = 96 (Cin = 32) -- parity, per-launch times, bench
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_r02o.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02o.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02o.txt 2>&1; grep wgrad gpurun_out/train_layers_r02o.txt | cut -c1-100; tail -1 gpurun_out/train_layers_r02o.txt | cut -c1-420
AESR_WGRAD_NO_FOLD=1 timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02o_nofold.txt 2>&1; tail -1 gpurun_out/train_layers_r02o_nofold.txt | cut -c1-420
timeout 600 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02o.json"))
t = d["train"]
print("infer", round(d["value"]), round(d["e2e"]["value"]), "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
