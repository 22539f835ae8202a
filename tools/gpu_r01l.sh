#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_r01l.log 2>&1; tail -6 gpurun_out/pytest_r01l.log
