#!/bin/bash
TAG=${1:-r2c16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for l in 1 0; do
  CP360_ROW_LIST=$l timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad' -f -o $OUT/row_l$l python tools/prof_one.py cubepad 64 128 1 0 32 > $OUT/row_l$l.log 2>&1; echo "ncu list=$l rc=$?"
  timeout 100 python tools/ncu_summary.py $OUT/row_l$l.ncu-rep > $OUT/row_l$l.txt 2>&1; grep -E "gpu__time|inst_executed|issue_active|bank_conflicts|stall|registers|dram__bytes" $OUT/row_l$l.txt | head -20
  timeout 100 python tools/ncu_lines.py $OUT/row_l$l.ncu-rep 16 > $OUT/row_l${l}_lines.txt 2>&1; head -18 $OUT/row_l${l}_lines.txt | cut -c1-150
  rm -f $OUT/row_l$l.ncu-rep
done
