#!/bin/bash
# c2e backward: walk order / split / grid A-B, then one ncu capture with source lines
TAG=${1:-r2c18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2e_backward or c2e_max_with or reproducible" > $OUT/pytest_c2e.log 2>&1; echo "pytest c2e rc=$?"; tail -3 $OUT/pytest_c2e.log
for cfg in "1 8" "0 8" "1 6" "1 4" "1 12"; do set -- $cfg
  echo "order=$1 split=$2"; CP360_C2E_BWD_ORDER=$1 CP360_C2E_BWD_SPLIT=$2 timeout 300 python tools/kbench.py --only bwd 2>&1 | grep -E "c2e bwd"
done | tee $OUT/kbench_cfg.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'c2e_bwd' -f -o $OUT/c2ebwd python tools/prof_one.py c2ebwd 8 1000 32 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/c2ebwd.ncu-rep > $OUT/c2ebwd.txt 2>&1; cat $OUT/c2ebwd.txt | head -30
timeout 100 python tools/ncu_lines.py $OUT/c2ebwd.ncu-rep 30 > $OUT/c2ebwd_lines.txt 2>&1; head -45 $OUT/c2ebwd_lines.txt | cut -c1-170
rm -f $OUT/c2ebwd.ncu-rep
