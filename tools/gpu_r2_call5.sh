#!/bin/bash
# Round 2, call 5: generate the tiling table; arrive-all vs product build speed; fused chain after the epilogue rework.
TAG=${1:-r2c5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -x -q -k "not builtin_table" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log; lap pytest
timeout 900 python tools/tune_table.py --out $OUT --rounds 2 --effort 2 > $OUT/tune_stdout.txt 2>&1; echo "tune rc=$?"; tail -3 $OUT/tune_stdout.txt; lap tune
CP360_AUTOTUNE=1 timeout 300 python tools/site_sweep.py --cube 224,256 --batch 32 > $OUT/sweep_product.txt 2>&1; echo "sweep product rc=$?"; lap sweep_product
CP360_AUTOTUNE=1 CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_arriveall.so timeout 300 python tools/site_sweep.py --cube 224,256 --batch 32 > $OUT/sweep_arriveall.txt 2>&1; echo "sweep arriveall rc=$?"; lap sweep_arriveall
paste -d'\n' $OUT/sweep_product.txt $OUT/sweep_arriveall.txt | grep -E "algo 6|total" | cut -c1-120
CP360_AUTOTUNE=1 CP360_BENCH_SITES=1 timeout 400 python bench.py --steps 100 --no-cpu-baseline --no-e2e --no-aten-baseline > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_256.json")); print(d["value"], d["ms_per_step"]); f=d["fused_chain"]; print(f["value"], f["ms_per_step"]); print(json.dumps(f["sites"]))
PY
lap bench
CP360_AUTOTUNE=1 CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_arriveall.so timeout 400 python bench.py --steps 100 --no-cpu-baseline --no-e2e --no-aten-baseline --no-fused > $OUT/bench_256_arriveall.json 2> $OUT/bench_256_arriveall.err; echo "bench 256 (arrive-all) rc=$?"; cut -c1-200 $OUT/bench_256_arriveall.json; lap bench_arriveall
timeout 300 python bench.py --workload clstm --no-cpu-baseline --no-e2e > $OUT/bench_clstm.json 2> $OUT/bench_clstm.err; echo "bench clstm rc=$?"; cut -c1-200 $OUT/bench_clstm.json; lap clstm
