#!/bin/bash
# round-1 evidence: launch list of the bench command, full ncu capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_s2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
F=2368 python scripts/profile_target.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:freg -s 2 -c 1 -f -o gpurun_out/r1_final_freg python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
F=2368 ncu --set full --clock-control none --import-source on -k regex:topm -s 2 -c 1 -f -o gpurun_out/r1_final_topm python scripts/profile_target.py > gpurun_out/ncu_full_topm.log 2>&1
echo "topm rc=$?"
tail -2 gpurun_out/bench_s2.log | cut -c1-300
