#!/bin/bash
TAG=${1:-r2c15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300; lap pytest
for l in 1 0; do echo "== CP360_ROW_LIST=$l"; CP360_ROW_LIST=$l timeout 300 python tools/site_sweep.py --cube 224,256 --batch 32 2>&1 | grep -E "algo 5|total"; done; lap sweep
for l in 1 0; do
  CP360_ROW_LIST=$l CP360_BENCH_SITES=1 timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_256_l$l.json 2> $OUT/bench_256_l$l.err; echo "bench list=$l rc=$?"; grep "cubepad_row" $OUT/bench_256_l$l.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_256_l$l.json")); f=d["fused_chain"]; print("list=$l", d["value"], d["ms_per_step"], "fused", f["value"])
PY
done; lap bench
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_row.log python tools/racecheck_probe.py --only row > $OUT/racecheck_row.out 2>&1; echo "racecheck row rc=$?"; tail -1 $OUT/racecheck_row.log )
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_row.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cubepad and not full_size and not resnet50_sites and not selftest" > $OUT/memcheck_row.out 2>&1; echo "memcheck cubepad rc=$?"; tail -1 $OUT/memcheck_row.log; tail -1 $OUT/memcheck_row.out; lap sanitizers
