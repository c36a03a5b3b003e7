#!/bin/bash
# round-2 evidence on one B200: shapes table, bench lines (headline, road-like generators, BASELINE configs 3 / 4), launch list,
# full ncu captures of the stream and select kernels for four shapes, per-call latency, get_lanes caller-level timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/shapes.py > gpurun_out/r2_shapes.log 2>&1; echo "shapes rc=$?"
# bench lines
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
for g in 2 3 4; do timeout 200 python bench.py --groups $g --no-e2e --no-cpu-baseline --no-ref-cuda --steps 30 > gpurun_out/r2_bench_groups$g.json 2>> gpurun_out/r2_bench.err; done
timeout 200 python bench.py --top-k 8 --no-e2e --no-cpu-baseline --no-ref-cuda --steps 30 > gpurun_out/r2_bench_topk8.json 2>> gpurun_out/r2_bench.err
# BASELINE config 3 (VIL-100: 240 priors x 36 offsets, top_k 8) and config 4 (sweep points)
timeout 200 python bench.py --proposals 240 --offsets 36 --top-k 8 --frames 65536 --e2e-frames 8192 --no-cpu-baseline --steps 30 > gpurun_out/r2_bench_config3_vil.json 2>> gpurun_out/r2_bench.err
: > gpurun_out/r2_bench_config4_sweep.jsonl
for n in 256 2048 8192; do for thr in 10 30 50; do
  timeout 200 python bench.py --proposals $n --overlap $thr --frames $((16384*1000/n)) --no-e2e --no-cpu-baseline --no-ref-cuda --steps 20 >> gpurun_out/r2_bench_config4_sweep.jsonl 2>> gpurun_out/r2_bench.err
done; done
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
rm -f gpurun_out/r2_bench_config4_sweep.jsonl.tmp
# launch list of the bench command
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_bench_s2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms -c 30 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
# (full ncu captures: scripts/gpu_r2_ncu.sh)
timeout 200 python scripts/bench_latency.py > gpurun_out/r2_latency.json 2>/dev/null; echo "latency rc=$?"
timeout 200 python scripts/bench_latency_detail.py 2>/dev/null | tail -1 > gpurun_out/r2_latency_detail.json
printf "1000 72 4 16384 8 0.1\n1000 72 4 16384 2 0.1\n1000 72 8 16384 8 0.1\n1000 36 8 16384 8 0.1\n240 72 4 32768 3 0.1\n240 36 8 32768 3 0.1\n4096 72 4 2048 8 0.1\n" | bash scripts/gpu_r2_split.sh > /dev/null
timeout 300 python scripts/bench_get_lanes.py > gpurun_out/r2_get_lanes_decode_bench.log 2>&1; echo "get_lanes rc=$?"
