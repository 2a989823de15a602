#!/bin/bash
# Round-end GPU-box pass at the bench default (32 frames per step), most important evidence first so that a
# call cut short still leaves the earlier files: parity tests, smoke, bench line (+ per-site timing), reference
# arm, ncu launch list, ncu --set full captures (dominant row-kernel sites first), per-kernel bench.
# Usage (from the repo root on the box): bash tools/gpu_final.sh <tag> [frames per step]
TAG=${1:-final}
B=${2:-32}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log; lap smoke
CP360_BENCH_SITES=1 timeout 400 python bench.py --batch $B > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -24 $OUT/bench.err; lap bench
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cat $OUT/bench_reference.json; lap reference
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --batch $B --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --profile-range > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; lap launches
for spec in "cubepad 64 128 1" "cubepad 128 64 1" "cubepad 64 64 1" "cubepad 3 256 3" "e2c 256" "c2emax 8 1000" "cubepad 256 32 1" "cubepad 128 32 1" \
            "cubepad 256 16 1" "cubepad 512 16 1" "cubepad 2048 8 1" "cubepad 512 8 1"; do
  name=$(echo $spec | tr ' ' '_')
  case "$spec" in cubepad*) args="$spec 0 $B";; *) args="$spec $B";; esac
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -f -o $OUT/full_$name \
      python tools/prof_one.py $args > $OUT/full_$name.log 2>&1; echo "ncu full $name rc=$?"; lap $name
done
for sec in bwd c2e e2c fused; do
  timeout 200 python tools/kbench.py --only $sec > $OUT/kbench_$sec.txt 2>&1; echo "kbench $sec rc=$?"; lap kbench_$sec
done
cat $OUT/kbench_*.txt > $OUT/kbench.txt
ls -la $OUT
