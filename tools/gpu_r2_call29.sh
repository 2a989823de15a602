#!/bin/bash
TAG=${1:-r2c29}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or bwd or grad" > $OUT/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"; tail -3 $OUT/pytest_bwd.log
timeout 300 python tools/experiments/cubepad_bwd_time.py | tee $OUT/bwd_B.txt
echo "cache off"; CP360_BWD_TABLE_CACHE=0 timeout 300 python tools/experiments/cubepad_bwd_time.py 256x32 | tee $OUT/bwd_B_nocache.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/racecheck_probe.py --only bwd > $OUT/${tool}_bwd.log 2>&1; echo "$tool rc=$?"; tail -5 $OUT/${tool}_bwd.log
done
timeout 900 python -m pytest tests/test_reference_callsites_gpu.py -m gpu -x -q > $OUT/pytest_ref.log 2>&1; echo "pytest ref rc=$?"; tail -2 $OUT/pytest_ref.log
