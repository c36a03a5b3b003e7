#!/bin/bash
# first GPU contact: fixtures from the live reference op, then the parity tests, under hard timeouts
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
timeout 600 python tests/golden/make_ref_fixtures.py gpurun_out/golden > gpurun_out/fixtures.log 2>&1; echo "fixtures rc=$?"
tail -3 gpurun_out/fixtures.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
