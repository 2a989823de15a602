#!/bin/bash
TAG=${1:-r2c10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300; lap pytest
timeout 300 python tools/kbench.py --only bwd --batch 16 > $OUT/kbench_bwd.txt 2>&1; echo "kbench bwd rc=$?"; grep "cubepad bwd" $OUT/kbench_bwd.txt; lap kbench_bwd
for cfg in "CP360_BWD_ALGO=1" "CP360_BWD_STAGE_KB=96 CP360_BWD_STAGES=2" "CP360_BWD_STAGE_KB=32 CP360_BWD_STAGES=4" "CP360_BWD_WARPS=8" "CP360_BWD_STAGE_KB=32 CP360_BWD_STAGES=3"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/kbench.py --only bwd --batch 16 2>&1 | grep "cubepad bwd"
done; lap bwd_variants
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_product.log python tools/racecheck_probe.py --only bwd > $OUT/racecheck_product.out 2>&1; echo "racecheck rc=$?"
  tail -1 $OUT/racecheck_product.out; tail -1 $OUT/racecheck_product.log ); lap racecheck
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_probe.log python tools/racecheck_probe.py --only bwd > $OUT/memcheck_probe.out 2>&1; echo "memcheck rc=$?"; tail -1 $OUT/memcheck_probe.log; lap memcheck
