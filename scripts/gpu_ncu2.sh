#!/bin/bash
TAG=${1:-prof}
mkdir -p gpurun_out
python scripts/profile_target.py > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:freg -s 2 -c 1 -f -o gpurun_out/${TAG} python scripts/profile_target.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:phnms -c 8 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_target.py > /dev/null 2>&1
grep -v "^==" gpurun_out/${TAG}_launches.csv | awk -F'","' 'NR>1{print $5, $NF}' | cut -c1-120 | tail -4
