#!/bin/bash
# chain step vs. where the site tensors sit: default bench flags (aten baseline etc. run first) and shifted outputs
TAG=${1:-r2c33}
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python - <<PY
import json
d=json.load(open("$OUT/$1.json")); print("$1: %.1f frames/s  %.4f ms  dominant %.3f chain %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["chain_frac"]))
PY
  grep -E "site cubepad_row (64x64|128x64|64x128|3x256)" $OUT/$1.err | cut -c1-100
  grep -E "ptrs site (1|2|5) " $OUT/$1.err | cut -c40-150; }
CP360_BENCH_SITES=2 timeout 300 python bench.py > $OUT/default.json 2> $OUT/default.err; show default
Q="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-aten-baseline --no-fused"
for off in 0 1 4 16 64 256 1024; do
  CP360_PIPE_OUT_OFF_KB=$off CP360_BENCH_SITES=2 timeout 300 python bench.py $Q > $OUT/out_$off.json 2> $OUT/out_$off.err; show out_$off
done
for off in 4 64 1024; do
  CP360_PIPE_IN_OFF_KB=$off CP360_BENCH_SITES=2 timeout 300 python bench.py $Q > $OUT/in_$off.json 2> $OUT/in_$off.err; show in_$off
done
