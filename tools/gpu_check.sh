#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, per-kernel bench, ncu launch list + full captures.
# Usage (from the repo root on the box): bash tools/gpu_check.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
CP360_BENCH_SITES=1 timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -30 $OUT/bench.err
timeout 600 python tools/kbench.py --json $OUT/kbench.json > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --profile-range > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for spec in "cubepad 3 256 3" "cubepad 64 128 1" "cubepad 64 64 1" "cubepad 128 64 1" "cubepad 128 32 1" "cubepad 256 32 1" "cubepad 256 16 1" "cubepad 512 16 1" "cubepad 512 8 1" "cubepad 2048 8 1" "e2c 256" "c2emax 8 1000"; do
  name=$(echo $spec | tr ' ' '_')
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -f -o $OUT/full_$name \
      python tools/prof_one.py $spec > $OUT/full_$name.log 2>&1; echo "ncu full $name rc=$?"
done
ls -la $OUT
