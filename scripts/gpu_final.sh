#!/bin/bash
# round-end evidence: tests, smoke, bench (both arms), launch list, full ncu captures
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -2
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 280 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.log
timeout 100 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>/dev/null; cut -c1-160 gpurun_out/bench_ref.log
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_s2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms -c 20 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
F=2368 python scripts/profile_target.py > gpurun_out/prof_plain.log 2>&1 &&
F=2368 ncu --set full --clock-control none --import-source on -k regex:freg -s 2 -c 1 -f -o gpurun_out/r1_final_freg python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
F=16384 ncu --set full --clock-control none --import-source on -k regex:topm -s 2 -c 1 -f -o gpurun_out/r1_final_topm python scripts/profile_target.py > gpurun_out/ncu_full_topm.log 2>&1
echo "topm rc=$?"
timeout 200 python scripts/shapes.py > gpurun_out/shapes.log 2>&1; cut -c1-200 gpurun_out/shapes.log
timeout 100 python scripts/bench_ref_cuda.py 2>/dev/null | tail -1 > gpurun_out/ref_cuda.json; cat gpurun_out/ref_cuda.json | cut -c1-400
