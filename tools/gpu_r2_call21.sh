#!/bin/bash
TAG=${1:-r2c21}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or bwd or grad" > $OUT/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"; tail -3 $OUT/pytest_bwd.log
for r in 1 0; do echo "CP360_BWD_REG_POS=$r"; CP360_BWD_REG_POS=$r timeout 300 python tools/kbench.py --only bwd 2>&1 | grep -E "cubepad bwd"; done | tee $OUT/kbench_bwd.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/racecheck_probe.py --only bwd > $OUT/${tool}_bwd.log 2>&1; echo "$tool rc=$?"; tail -5 $OUT/${tool}_bwd.log
done
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad_bwd' -f -o $OUT/cpbwd python tools/prof_one.py cubepadbwd 256 32 1 32 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/cpbwd.ncu-rep > $OUT/cpbwd.txt 2>&1; cat $OUT/cpbwd.txt | head -30
timeout 100 python tools/ncu_lines.py $OUT/cpbwd.ncu-rep 30 > $OUT/cpbwd_lines.txt 2>&1; head -20 $OUT/cpbwd_lines.txt | cut -c1-170
rm -f $OUT/cpbwd.ncu-rep
