#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep.py > gpurun_out/sweep_N1000_No72.log 2>&1; tail -30 gpurun_out/sweep_N1000_No72.log
