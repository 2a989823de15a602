#!/bin/bash
# Regenerate the CubePad tiling table on a B200: isolated pass (tune_table.py), then the in-chain pass (tune_chain.py);
# rebuild with the new table on the box and check it (tests + bench lines at both face widths, B = 1 and 32).
TAG=${1:-r2tune}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python tools/tune_table.py --out $OUT/iso --rounds 3 --effort 2 > $OUT/tune_table_stdout.txt 2>&1; echo "tune_table rc=$?"; tail -2 $OUT/tune_table_stdout.txt; lap iso
timeout 1800 python tools/tune_chain.py --base $OUT/iso/cubepad_tuned.h --out $OUT > $OUT/tune_chain_stdout.txt 2>&1; echo "tune_chain rc=$?"; grep "pass 2\|clstm" $OUT/tune_chain_stdout.txt | tail -40; lap chain
cp $OUT/cubepad_tuned.h cp-360-weakly-supervised-saliency_b200/csrc/cubepad_tuned.h
timeout 600 python -c "import cp360_b200; print(cp360_b200.build_library(force=True))"; lap rebuild
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300; lap pytest
for cube in 256 224; do for b in 1 8 32; do
  CP360_BENCH_SITES=1 timeout 200 python bench.py --cube $cube --batch $b --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-aten-baseline > $OUT/bench_${cube}_b$b.json 2> $OUT/bench_${cube}_b$b.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${cube}_b$b.json")); f=d.get("fused_chain") or {}
    print("cube $cube B %2d: %9.1f frames/s %8.4f ms/step chain %6.1f GB/s (%.3f) dominant %s %.3f fused %s frames/s" % ($b, d["value"], d["ms_per_step"], d["roofline"]["chain_gbs"], d["roofline"]["chain_frac"], d["roofline"]["kernel"], d["roofline"]["frac"], f.get("value")))
except Exception as e: print("cube $cube B $b: no line", e)
PY
done; done; lap bench
cat $OUT/bench_256_b32.err $OUT/bench_224_b32.err
