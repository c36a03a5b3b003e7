#!/bin/bash
# per-kernel split (ncu launch list, durations only) of the op for a list of shapes: "N NOFF TOPK F GROUPS OUTL" per line
cd "$(dirname "$0")/.."
OUT=${OUT:-gpurun_out/r2_split.log}
: > $OUT
while read -r N NOFF TOPK F GRP OUTL; do
  [ -z "$N" ] && continue
  tag="N${N}_No${NOFF}_k${TOPK}_g${GRP}_o${OUTL}"
  N=$N NOFF=$NOFF TOPK=$TOPK F=$F NGROUPS=$GRP OUTL=$OUTL REPS=3 python scripts/profile_target.py > /dev/null 2>&1 || { echo "$tag plain run failed" >> $OUT; continue; }
  N=$N NOFF=$NOFF TOPK=$TOPK F=$F NGROUPS=$GRP OUTL=$OUTL REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms --csv --log-file gpurun_out/_split.csv python scripts/profile_target.py > /dev/null 2>&1
  python - "$tag" >> $OUT <<'P'
import csv, sys, collections
rows = [r for r in csv.reader(open('gpurun_out/_split.csv')) if len(r) > 5]
h = rows[0]; ik, iv = h.index('Kernel Name'), h.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ik].split('(')[0].replace('void ', '').replace('phnms::', ''), []).append(float(r[iv].replace(',', '')))
print(sys.argv[1], '  '.join(f"{k}: {v[-1] / 1000:.1f} us" for k, v in d.items()))
P
done
rm -f gpurun_out/_split.csv
cat $OUT
