#!/bin/bash
cd "$(dirname "$0")/.."
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_t9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_t9.log
tail -6 gpurun_out/r2_t9.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_bench1.json
timeout 300 python bench.py --groups 3 --no-e2e --no-cpu-baseline --no-ref-cuda --steps 30 > gpurun_out/r2_bench_g3.json 2>> gpurun_out/r2_bench1.err; cut -c1-300 gpurun_out/r2_bench_g3.json
