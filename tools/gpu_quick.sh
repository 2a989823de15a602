#!/bin/bash
# Short GPU-box pass: parity tests, smoke, bench, the e2c / c2e / fused sections of kbench.
# Usage (from the repo root on the box): bash tools/gpu_quick.sh <tag>
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
CP360_BENCH_SITES=1 timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -30 $OUT/bench.err
for sec in e2c c2e; do
  timeout 300 python tools/kbench.py --only $sec > $OUT/kbench_$sec.txt 2>&1; echo "kbench $sec rc=$?"; cat $OUT/kbench_$sec.txt
done
