#!/bin/bash
# rows per step of the wide-row copy loop: EPI 2 (product) / 3 / 4, plain 1 (product) / 2
TAG=${1:-r2c34}
OUT=gpurun_out/$TAG
mkdir -p $OUT
L=$PWD/cp-360-weakly-supervised-saliency_b200/lib
for v in ${VARIANTS:-product epi3 epi4 wide2 product}; do
  if [ $v = product ]; then unset CP360_LIB; else export CP360_LIB=$L/libcp360_$v.so; fi
  CP360_BENCH_SITES=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-aten-baseline > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$v.json")); f=d.get("fused_chain") or {}
print("$v: %.1f frames/s  %.4f ms  dominant %.3f chain %.3f | fused %.1f frames/s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["chain_frac"], f.get("value", 0)))
fs=f.get("sites") or {}
for k in fs:
    if "64x128" in k or "3x256" in k: print("   fused site", k, fs[k])
PY
  grep -E "site cubepad_row (64x128|3x256)" $OUT/bench_$v.err | cut -c1-100
done
