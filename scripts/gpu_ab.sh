#!/bin/bash
# A/B of kernel builds (PHNMS_SO) on the main shapes + parity tests of the default build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
S="1000x72x4x16384,1000x72x8x8192,1000x36x4x16384,1000x36x8x16384,240x72x4x32768"
for so in ${AB_BUILDS:-ab_head ab_pk1 libphnms}; do
  echo "== $so"
  PHNMS_SO=$PWD/phnet_b200/csrc/$so.so SHAPES=$S timeout 200 python scripts/shapes.py 2>&1 | tee gpurun_out/shapes_$so.log | cut -c1-60,150-230
done
VARIANTS=2 timeout 150 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log | cut -c1-200
