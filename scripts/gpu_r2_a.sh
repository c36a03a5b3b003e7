#!/bin/bash
# round 2, run A: stream-path tests to completion, launch list of three shapes, a few tuning variants
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_stream_gpu.py -q > gpurun_out/r2_t2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_t2.log
tail -4 gpurun_out/r2_t2.log
export SHAPES=1000x72x4x4096x8,1000x72x4x4096x2,1000x72x8x4096x8,240x36x8x16384x8
timeout -k 10 200 python scripts/shapes.py > gpurun_out/r2_sh_plain.log 2>&1 && \
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms --csv --log-file gpurun_out/r2_launches_a.csv python scripts/shapes.py > gpurun_out/r2_sh_ncu.log 2>&1
unset SHAPES
SHAPES=1000x72x4x16384x8,1000x72x4x16384x2,1000x72x8x16384x8,1000x36x8x16384x8 TUNES='[null,{"lanes_per_pass":2},{"stream_warps":14},{"stream_warps":12},{"stream_warps":8},{"path":1,"variant":2}]' timeout -k 10 400 python scripts/shapes.py > gpurun_out/r2_shapes2.log 2>&1
cat gpurun_out/r2_shapes2.log
