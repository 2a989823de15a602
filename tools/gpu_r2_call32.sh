#!/bin/bash
# run-to-run spread of the chain on ONE box: same command five times, per-site times and tensor placement
TAG=${1:-r2c32}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for i in 1 2 3 4 5; do
  steps=200; [ $i -ge 4 ] && steps=100
  CP360_BENCH_SITES=2 timeout 300 python bench.py --steps $steps --warmup 10 --no-cpu-baseline --no-e2e --no-aten-baseline --no-fused > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$i.json")); print("run $i (steps $steps): %.1f frames/s  %.4f ms  dominant %.3f chain %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["chain_frac"]))
PY
  grep -E "site cubepad_row (64x64|128x64|64x128)" $OUT/bench_$i.err | cut -c1-100
  grep -E "ptrs site (1|2|3|4|5) " $OUT/bench_$i.err | cut -c1-150
done
