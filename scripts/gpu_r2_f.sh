#!/bin/bash
cd "$(dirname "$0")/.."
export SHAPES=1000x72x4x16384x8,1000x72x4x16384x2,240x36x8x32768x8
for v in "" _k8c6 _k16c5 _k16c4 _k32c5; do
  echo "== select only, variant libphnms$v"
  PHNMS_SO=phnet_b200/csrc/libphnms$v.so PHNMS_SKIP=6 timeout -k 10 120 python scripts/shapes.py 2>&1 | cut -c1-60,200-240
done
unset SHAPES
export F=16384
python scripts/profile_target.py > gpurun_out/r2_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:select -s 2 -c 1 -f -o gpurun_out/r2_select_v2 python scripts/profile_target.py > gpurun_out/r2_ncu_select2.log 2>&1
echo "select ncu rc=$?"
