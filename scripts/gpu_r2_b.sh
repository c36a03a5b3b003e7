#!/bin/bash
# round 2, run B: failing tests with step reporting, then full ncu captures of the stream and select kernels
cd "$(dirname "$0")/.."
PHNMS_DEBUG=1 timeout -k 10 600 python -m pytest tests/test_stream_gpu.py -q -k "draw_cap or ragged" > gpurun_out/r2_t3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_t3.log
grep -E "phnms:|passed|failed" gpurun_out/r2_t3.log | sort | uniq -c | head
export F=4736 TUNE='{"lanes_per_pass":2}'
python scripts/profile_target.py > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stream -s 2 -c 1 -f -o gpurun_out/r2_stream_l2 python scripts/profile_target.py > gpurun_out/r2_ncu_stream.log 2>&1
echo "stream ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:select -s 2 -c 1 -f -o gpurun_out/r2_select python scripts/profile_target.py > gpurun_out/r2_ncu_select.log 2>&1
echo "select ncu rc=$?"
