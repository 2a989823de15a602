#!/bin/bash
# per-line profile of the row kernel at the W = 64 sites
TAG=${1:-r2c30}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in cubepad_64_64_1 cubepad_128_64_1; do
  CP360_PROF_CUBE=256 CP360_PROF_ONLY=$spec timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad' -f \
      -o $OUT/src_$spec python tools/prof_all.py 32 > $OUT/src_$spec.log 2>&1; echo "ncu $spec rc=$?"
  timeout 120 python tools/ncu_summary.py $OUT/src_$spec.ncu-rep 2>&1 | grep -E "time_duration|inst_executed.sum|issue_active|stall" | head -8
  timeout 120 python tools/ncu_lines.py $OUT/src_$spec.ncu-rep 50 > $OUT/lines_$spec.txt 2>&1; rm -f $OUT/src_$spec.ncu-rep
done
head -40 $OUT/lines_cubepad_64_64_1.txt | cut -c1-160
