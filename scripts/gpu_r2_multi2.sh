#!/bin/bash
# 2 GPUs: peer-memory collection end to end (tests/peer_check.py), the 2-GPU test, bench with both collection methods
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29631 tests/peer_check.py > gpurun_out/r2_peer_check_2gpu.log 2>&1; echo "peer_check rc=$?"; tail -3 gpurun_out/r2_peer_check_2gpu.log
timeout 600 python -m pytest tests/test_collect_gpu.py -q 2>&1 | tail -2
for mode in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 30 --warmup 5 --collect $mode > gpurun_out/r2_bench_2gpu_$mode.json 2> gpurun_out/r2_bench_2gpu_$mode.err; echo "bench $mode rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_2gpu_$mode.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['collection_verified'], d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], d['e2e']['h2d_ceiling_gbs_per_gpu'], d['config']['per_rank'])"
done
