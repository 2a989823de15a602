#!/bin/bash
# Round 2, call 6: new c2e kernels (cluster K3m, gather backward), per-thread arrive under racecheck, table-driven tiling.
TAG=${1:-r2c6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log; lap smoke
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_product.log python tools/racecheck_probe.py > $OUT/racecheck_product.out 2>&1; echo "racecheck (product build, 2-stage rings, all kernels) rc=$?"
  tail -2 $OUT/racecheck_product.out; tail -2 $OUT/racecheck_product.log ); lap racecheck
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_probe.log python tools/racecheck_probe.py > $OUT/memcheck_probe.out 2>&1; echo "memcheck probe rc=$?"; tail -2 $OUT/memcheck_probe.log; lap memcheck
timeout 300 python tools/kbench.py --only c2e --batch 32 > $OUT/kbench_c2e.txt 2>&1; echo "kbench c2e rc=$?"; cat $OUT/kbench_c2e.txt; lap kbench_c2e
timeout 300 python tools/kbench.py --only bwd --batch 16 > $OUT/kbench_bwd.txt 2>&1; echo "kbench bwd rc=$?"; cat $OUT/kbench_bwd.txt; lap kbench_bwd
CP360_BENCH_SITES=1 timeout 400 python bench.py --steps 100 --no-cpu-baseline --no-aten-baseline > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; tail -14 $OUT/bench_256.err
python - <<PY
import json
d=json.load(open("$OUT/bench_256.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); f=d["fused_chain"]; print(f["value"], f["ms_per_step"]); print(json.dumps(d["tuning"]))
PY
lap bench
for cfg in "CP360_ROW_CTAS=2" "CP360_ROW_CTAS=2 CP360_ROW_SLOTS=2"; do
  env $cfg timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_tmp.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("$OUT/bench_tmp.json")); f=d["fused_chain"]; print("$cfg", d["value"], f["value"], {k:v["us"] for k,v in f["sites"].items()})
PY
done; lap rowctas
