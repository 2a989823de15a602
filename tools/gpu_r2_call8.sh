#!/bin/bash
TAG=${1:-r2c8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 100 python tools/experiments/debug_nan.py 2>&1 | tail -8; lap debug
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|NaN pattern" $OUT/pytest_gpu.log | cut -c1-400; lap pytest
timeout 300 python tools/kbench.py --only c2e --batch 32 > $OUT/kbench_c2e.txt 2>&1; echo "kbench c2e rc=$?"; grep -E "\[192,(1000|2048),[78]" $OUT/kbench_c2e.txt; lap kbench_c2e
timeout 300 python tools/kbench.py --only bwd --batch 16 > $OUT/kbench_bwd.txt 2>&1; echo "kbench bwd rc=$?"; grep -v "torch copy" $OUT/kbench_bwd.txt; lap kbench_bwd
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_product.log python tools/racecheck_probe.py > $OUT/racecheck_product.out 2>&1; echo "racecheck rc=$?"
  tail -1 $OUT/racecheck_product.out; tail -1 $OUT/racecheck_product.log ); lap racecheck
CP360_BENCH_SITES=1 timeout 400 python bench.py --steps 100 --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; tail -14 $OUT/bench_256.err
python - <<PY
import json
d=json.load(open("$OUT/bench_256.json")); print(d["value"], d["ms_per_step"]); f=d["fused_chain"]; print(f["value"], f["ms_per_step"], {k:v["us"] for k,v in f["sites"].items()})
PY
lap bench
