#!/bin/bash
# usage: gpu_ncu.sh <tag> ; env CLUSTER THREADS F etc. are forwarded
set -x
TAG=${1:-prof}
mkdir -p gpurun_out
python scripts/profile_target.py > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-freg|fused_kernel} -s ${SKIP:-2} -c 1 -f -o gpurun_out/${TAG} python scripts/profile_target.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_plain.log; tail -5 gpurun_out/${TAG}_ncu.log
