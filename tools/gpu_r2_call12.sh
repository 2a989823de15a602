#!/bin/bash
TAG=${1:-r2c12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300; lap pytest
timeout 300 python tools/kbench.py --only c2e --batch 32 > $OUT/kbench_c2e.txt 2>&1; echo "kbench c2e rc=$?"; grep -E "\[192,(1000|2048),[78]" $OUT/kbench_c2e.txt; lap kbench
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_probe.log python tools/racecheck_probe.py > $OUT/synccheck_probe.out 2>&1; echo "synccheck probe rc=$?"; tail -1 $OUT/synccheck_probe.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_suite.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model" > $OUT/synccheck_suite.out 2>&1; echo "synccheck suite rc=$?"; tail -1 $OUT/synccheck_suite.log; tail -2 $OUT/synccheck_suite.out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_c2e.log python tools/racecheck_probe.py --only c2e > $OUT/racecheck_c2e.out 2>&1; echo "racecheck c2e rc=$?"; tail -1 $OUT/racecheck_c2e.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_c2e.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c2e" > $OUT/memcheck_c2e.out 2>&1; echo "memcheck c2e rc=$?"; tail -1 $OUT/memcheck_c2e.log; lap sanitizers
CP360_BENCH_SITES=1 timeout 500 python bench.py > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; cut -c1-200 $OUT/bench_256.json; tail -13 $OUT/bench_256.err
python - <<PY
import json
d=json.load(open("$OUT/bench_256.json")); e=d["e2e"]; print(d["value"], e["value"], e["h2d_gbs_aggregate"], e["h2d_ceiling_gbs_aggregate"], e["frac_of_h2d_ceiling"], d["fused_chain"]["value"], d["gpu_aten_baseline"]["value"], d["cpu_baseline"]["value"])
PY
lap bench
