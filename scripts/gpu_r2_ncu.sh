#!/bin/bash
# full ncu captures of the stream / select kernels (one B200)
cd "$(dirname "$0")/.."
PART=${1:-all}
cap() { # name, kernel regex, env...
  name=$1; shift; kre=$1; shift
  env "$@" python scripts/profile_target.py > gpurun_out/r2_prof_plain_$name.log 2>&1 &&
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$kre -s 2 -c 1 -f -o gpurun_out/r2_$name python scripts/profile_target.py > gpurun_out/r2_ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
# (each report is ~11 MB and one call may bring back 64 MB: two parts)
if [ "$PART" = "a" ] || [ "$PART" = "all" ]; then
cap stream_72_k4 stream F=16384 N=1000 NOFF=72 TOPK=4
cap select_72_k4 select F=16384 N=1000 NOFF=72 TOPK=4
cap stream_72_k8 stream F=16384 N=1000 NOFF=72 TOPK=8
fi
if [ "$PART" = "b" ] || [ "$PART" = "all" ]; then
cap stream_36_k8 stream F=16384 N=1000 NOFF=36 TOPK=8
cap stream_4096 stream F=2048 N=4096 NOFF=72 TOPK=4
cap stream_72_g2 stream F=16384 N=1000 NOFF=72 TOPK=4 NGROUPS=2
fi
if [ "$PART" = "c" ] || [ "$PART" = "all" ]; then      # the one-launch small-frame kernel on PHNet's own shapes
cap small_240_72_k4 small F=32768 N=240 NOFF=72 TOPK=4 NGROUPS=3
cap small_240_36_k8 small F=32768 N=240 NOFF=36 TOPK=8 NGROUPS=3
fi
# one-frame calls: device time of the single launch
F=1 N=240 NOFF=72 TOPK=4 REPS=20 python scripts/profile_target.py > /dev/null 2>&1 &&
F=1 N=240 NOFF=72 TOPK=4 REPS=20 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms --csv --log-file gpurun_out/r2_one_frame_launches.csv python scripts/profile_target.py > /dev/null 2>&1
echo "one-frame rc=$?"
