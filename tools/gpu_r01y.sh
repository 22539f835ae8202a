#!/bin/bash
# r01y: cached synthesis plans (no pageable H2D index copies per call) + synthesized-only download: full suite + e2e
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01y.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r01y.log
for g in 2 1 4; do
timeout 400 python bench.py --steps 20 --no-train --cpu-sample 1 --groups $g > gpurun_out/bench_r01y_g$g.json 2> gpurun_out/bench_r01y_g$g.err; echo "bench g$g rc $?"
done
