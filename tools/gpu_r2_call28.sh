#!/bin/bash
TAG=${1:-r2c28}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "2000 7 1 1" "2000 7 1 16" "256 32 1 1" "256 32 1 16"; do
  set -- $cfg
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad_bwd' -f -o $OUT/b_$1_$2_$4 python tools/prof_one.py cubepadbwd $cfg > $OUT/ncu_$1_$2_$4.log 2>&1
  echo "== $cfg"; timeout 100 python tools/ncu_summary.py $OUT/b_$1_$2_$4.ncu-rep 2>&1 | grep -E "time_duration|inst_executed.sum|issue_active|grid_size|stall" | head -12
  timeout 100 python tools/ncu_lines.py $OUT/b_$1_$2_$4.ncu-rep 40 > $OUT/lines_$1_$2_$4.txt 2>&1
  rm -f $OUT/b_$1_$2_$4.ncu-rep
done
