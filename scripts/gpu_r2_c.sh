#!/bin/bash
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_stream_gpu.py -q > gpurun_out/r2_t4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_t4.log
tail -3 gpurun_out/r2_t4.log
export SHAPES=1000x72x4x16384x8,1000x72x4x16384x2,1000x72x8x16384x8,240x36x8x32768x8 TUNE='{"lanes_per_pass":2}'
for sk in 0 6 5 3; do echo "== PHNMS_SKIP=$sk (1 select, 2 stream, 4 resume skipped)"; PHNMS_SKIP=$sk timeout -k 10 200 python scripts/shapes.py 2>&1 | cut -c1-60,200-330; done > gpurun_out/r2_skip.log 2>&1
cat gpurun_out/r2_skip.log
