#!/bin/bash
# quick loop: fused-path parity tests, the tuning sweep, the phase trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nms_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
VARIANTS=${VARIANTS:-2} timeout 150 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; grep -E "threads\": (512|256)|BEST" gpurun_out/sweep.log | head -8
timeout 100 python scripts/trace_phases.py 2>&1 | tail -16
