#!/bin/bash
# Round 2, call 7: lane=channel c2e kernels, launch-bounds fix, EPI producer protocol; experiments for small sites and CubePad bwd.
TAG=${1:-r2c7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|NaN pattern" $OUT/pytest_gpu.log | cut -c1-400; lap pytest
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_product.log python tools/racecheck_probe.py > $OUT/racecheck_product.out 2>&1; echo "racecheck (product build, 2-stage rings, all kernels) rc=$?"
  tail -2 $OUT/racecheck_product.out; tail -2 $OUT/racecheck_product.log ); lap racecheck
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_probe.log python tools/racecheck_probe.py > $OUT/memcheck_probe.out 2>&1; echo "memcheck probe rc=$?"; tail -2 $OUT/memcheck_probe.log; lap memcheck
timeout 300 python tools/kbench.py --only c2e --batch 32 > $OUT/kbench_c2e.txt 2>&1; echo "kbench c2e rc=$?"; grep -E "\[192,(1000|2048),[78]" $OUT/kbench_c2e.txt; lap kbench_c2e
timeout 300 python tools/kbench.py --only bwd --batch 16 > $OUT/kbench_bwd.txt 2>&1; echo "kbench bwd rc=$?"; grep -v "torch copy" $OUT/kbench_bwd.txt; lap kbench_bwd
CP360_BENCH_SITES=1 timeout 400 python bench.py --steps 100 --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; tail -14 $OUT/bench_256.err
python - <<PY
import json
d=json.load(open("$OUT/bench_256.json")); print(d["value"], d["ms_per_step"]); f=d["fused_chain"]; print(f["value"], f["ms_per_step"], {k:v["us"] for k,v in f["sites"].items()})
PY
lap bench
for cfg in "CP360_TUNED_TABLE=0" "CP360_TUNED_TABLE=0 CP360_CUBE_CTAS=2 CP360_CUBE_STAGE_KB=24 CP360_CUBE_STAGES=3" "CP360_TUNED_TABLE=0 CP360_CUBE_CTAS=2 CP360_CUBE_STAGE_KB=24 CP360_CUBE_STAGES=4 CP360_CUBE_WARPS=8" "CP360_TUNED_TABLE=0 CP360_CUBE_CTAS=3 CP360_CUBE_STAGE_KB=16 CP360_CUBE_STAGES=3 CP360_CUBE_WARPS=8"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/site_sweep.py --cube 224 --batch 32 2>&1 | grep -E "256x 14|512x  7|512x 14|128x 28"
done; lap cube_ctas
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad' -f -o $OUT/full_cubepadbwd_256_32 python tools/prof_one.py cubepadbwd 256 32 1 16 > $OUT/full_cubepadbwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/full_cubepadbwd_256_32.ncu-rep > $OUT/ncu_cubepadbwd_256_32.txt 2>&1; cat $OUT/ncu_cubepadbwd_256_32.txt | head -40
timeout 100 python tools/ncu_lines.py $OUT/full_cubepadbwd_256_32.ncu-rep > $OUT/ncu_cubepadbwd_256_32_lines.txt 2>&1; head -40 $OUT/ncu_cubepadbwd_256_32_lines.txt
rm -f $OUT/*.ncu-rep; lap ncu
