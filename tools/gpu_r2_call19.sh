#!/bin/bash
TAG=${1:-r2c19}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2e_backward or c2e_max_with or reproducible" > $OUT/pytest_c2e.log 2>&1; echo "pytest c2e rc=$?"; tail -3 $OUT/pytest_c2e.log
timeout 300 python tools/experiments/c2e_bwd_time.py | tee $OUT/bwd_time.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'c2e_bwd' -f -o $OUT/c2ebwd python tools/prof_one.py c2ebwd 8 1000 32 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/c2ebwd.ncu-rep > $OUT/c2ebwd.txt 2>&1; cat $OUT/c2ebwd.txt | head -30
timeout 100 python tools/ncu_lines.py $OUT/c2ebwd.ncu-rep 30 > $OUT/c2ebwd_lines.txt 2>&1; head -16 $OUT/c2ebwd_lines.txt | cut -c1-170
rm -f $OUT/c2ebwd.ncu-rep
