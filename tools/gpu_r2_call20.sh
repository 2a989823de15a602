#!/bin/bash
TAG=${1:-r2c20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cubic or cv2" > $OUT/pytest_cubic.log 2>&1; echo "pytest cubic rc=$?"; tail -3 $OUT/pytest_cubic.log
for k in 0 32 16 8; do echo "CP360_CUBIC_K=$k"; if [ $k = 0 ]; then unset CP360_CUBIC_K; else export CP360_CUBIC_K=$k; fi
  timeout 300 python tools/kbench.py --only c2e 2>&1 | grep -E "cubic"; done | tee $OUT/kbench_cubic.txt
unset CP360_CUBIC_K
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubic' -f -o $OUT/cubic python tools/prof_one.py c2ecubic 8 1000 32 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/cubic.ncu-rep > $OUT/cubic.txt 2>&1; cat $OUT/cubic.txt | head -30
timeout 100 python tools/ncu_lines.py $OUT/cubic.ncu-rep 30 > $OUT/cubic_lines.txt 2>&1; head -24 $OUT/cubic_lines.txt | cut -c1-170
rm -f $OUT/cubic.ncu-rep
