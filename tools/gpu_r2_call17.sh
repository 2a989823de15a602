#!/bin/bash
# c2e backward with team-split slots: parity, timing (old G vs one-wave G), sanitizers
TAG=${1:-r2c17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2e" > $OUT/pytest_c2e.log 2>&1; echo "pytest c2e rc=$?"; tail -5 $OUT/pytest_c2e.log
timeout 300 python tools/kbench.py --only bwd 2>&1 | grep -E "c2e" | tee $OUT/kbench_bwd.txt
for g in 4 7 9 14; do echo "G=$g"; CP360_C2E_BWD_G=$g timeout 300 python tools/kbench.py --only bwd 2>&1 | grep -E "c2e bwd"; done | tee $OUT/kbench_bwd_G.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/racecheck_probe.py --only c2ebwd > $OUT/${tool}_c2ebwd.log 2>&1; echo "$tool rc=$?"; tail -4 $OUT/${tool}_c2ebwd.log
done
