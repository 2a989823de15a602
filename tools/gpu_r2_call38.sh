#!/bin/bash
# in-chain re-ranking of the CubePad tilings inside the 20-launch chain (first site fused into e2c), B = 16 / 32 / 64
TAG=${1:-r2c38}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python bench.py --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_before.json 2>/dev/null; python -c "import json; d=json.load(open('$OUT/bench_before.json')); print('before: %.1f frames/s' % d['value'])"; lap before
timeout 1500 python tools/tune_chain.py --base cp-360-weakly-supervised-saliency_b200/csrc/cubepad_tuned.h --out $OUT --frames ${FRAMES:-32,16,64} --skip-clstm > $OUT/tune_chain_stdout.txt 2>&1; echo "tune_chain rc=$?"; grep "pass\|start" $OUT/tune_chain_stdout.txt | tail -24; lap chain
cp $OUT/cubepad_tuned.h cp-360-weakly-supervised-saliency_b200/csrc/cubepad_tuned.h
timeout 600 python -c "import cp360_b200; print(cp360_b200.build_library(force=True))" | tail -1; lap rebuild
for i in 1 2; do for cube in 256 224; do
  timeout 300 python bench.py --cube $cube --no-cpu-baseline --no-aten-baseline --no-e2e > $OUT/bench_after_${cube}_$i.json 2>/dev/null; python -c "import json; d=json.load(open('$OUT/bench_after_${cube}_$i.json')); print('after cube $cube: %.1f frames/s chain %.3f' % (d['value'], d['roofline']['chain_frac']))"
done; done; lap after
timeout 600 python -m pytest tests -m gpu -q -x -k "cubepad or pipeline" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/pytest_gpu.log; lap pytest
