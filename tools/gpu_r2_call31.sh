#!/bin/bash
# row kernel with 8 (product) / 12 / 16 warps per CTA: chain step and per-site times
TAG=${1:-r2c31}
OUT=gpurun_out/$TAG
mkdir -p $OUT
L=$PWD/cp-360-weakly-supervised-saliency_b200/lib
for v in product w12 w16; do
  if [ $v = product ]; then unset CP360_LIB; else export CP360_LIB=$L/libcp360_$v.so; fi
  for cube in 256 224; do
    CP360_BENCH_SITES=1 timeout 300 python bench.py --cube $cube --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-aten-baseline --no-fused > $OUT/bench_${v}_$cube.json 2> $OUT/bench_${v}_$cube.err
    python - <<PY
import json
d=json.load(open("$OUT/bench_${v}_$cube.json")); print("$v cube $cube: %.1f frames/s  %.4f ms  dominant %.3f chain %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["chain_frac"]))
PY
    grep -E "row" $OUT/bench_${v}_$cube.err | cut -c1-150 | head -8
  done
done
