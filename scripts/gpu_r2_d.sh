#!/bin/bash
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_stream_gpu.py tests/test_nms_gpu.py -q -x > gpurun_out/r2_t6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_t6.log
tail -3 gpurun_out/r2_t6.log
export SHAPES=1000x72x4x16384x8,1000x72x4x16384x2,1000x72x8x16384x8,1000x36x4x16384x8,1000x36x8x16384x8,240x36x8x32768x8,240x72x4x32768x3
TUNES='[null,{"lanes_per_pass":4}]' timeout -k 10 300 python scripts/shapes.py > gpurun_out/r2_shapes3.log 2>&1
cut -c1-70,190-400 gpurun_out/r2_shapes3.log
echo "== CPT=1 for n_off 36"
SHAPES=1000x36x4x16384x8,1000x36x8x16384x8,240x36x8x32768x8 PHNMS_STREAM_CPT=1 TUNES='[null,{"lanes_per_pass":4}]' timeout -k 10 300 python scripts/shapes.py 2>&1 | cut -c1-70,190-400
echo "== select only"
PHNMS_SKIP=6 timeout -k 10 200 python scripts/shapes.py 2>&1 | cut -c1-70,190-250
