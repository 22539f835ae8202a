#!/bin/bash
# r02k: CUDA-graph replay of the training step -- training parity, bench, per-launch times
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_r02k.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02k.log
timeout 300 python tools/train_layer_times.py > gpurun_out/train_layers_r02k.txt 2>&1; echo "layers rc $?"; tail -3 gpurun_out/train_layers_r02k.txt | cut -c1-400
timeout 600 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02k.json"))
t = d["train"]
print("infer", round(d["value"]), "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
print(t["roofline"]["kernel_ms"])
PY
