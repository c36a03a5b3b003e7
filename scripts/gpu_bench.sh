#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python scripts/sweep.py > gpurun_out/sweep_N1000_No72.log 2>&1; tail -20 gpurun_out/sweep_N1000_No72.log
timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
