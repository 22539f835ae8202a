#!/bin/bash
# final check of the round: full GPU suite, smoke, default bench of both arms
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc $?"; head -c 300 gpurun_out/bench_final_ref.json; echo
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
t = d["train"]
print("infer", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], "train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
