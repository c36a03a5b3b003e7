#!/bin/bash
# 8 GPUs: bench with the peer-memory collection (verified against an NCCL all-gather on every rank) and with the NCCL all-gather
cd "$(dirname "$0")/.."
nvidia-smi -L | wc -l
for mode in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 30 --warmup 5 --collect $mode > gpurun_out/r2_bench_8gpu_$mode.json 2> gpurun_out/r2_bench_8gpu_$mode.err; echo "bench $mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_8gpu_$mode.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['collection_verified'], d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], d['e2e']['h2d_ceiling_gbs_per_gpu'], d['config']['per_rank'])"
done
