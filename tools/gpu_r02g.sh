#!/bin/bash
# r02g: 16-byte vectorised LPIPS head / max-pool backward / BN backward -- training parity + training bench
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_r02g.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_r02g.log
timeout 600 python bench.py --steps 10 --cpu-sample 1 > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02g.json"))
t = d["train"]
print("infer", round(d["value"]), "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
print(t["roofline"]["kernel_ms"])
PY
