#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01e.log 2>&1; tail -3 gpurun_out/pytest_r01e.log
python bench.py --steps 20 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -c 600 gpurun_out/bench_r01e.err
AESR_LINFOLD=0 python bench.py --steps 10 --no-train --cpu-sample 1 > gpurun_out/bench_r01e_nofold.json 2>&1
python tools/train_graph_probe.py > gpurun_out/train_graph_probe.log 2>&1; tail -5 gpurun_out/train_graph_probe.log
python tools/layer_times.py --enc 640 --dec 3456 > gpurun_out/layers_r01e_big.log 2>&1
