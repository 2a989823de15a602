#!/bin/bash
# Round 2, call 1 (existing code): compute-sanitizer passes (bounded) + baseline numbers for the 224 family and
# the small-batch regime. Usage: bash tools/gpu_r2_call1.sh <tag>
TAG=${1:-r2c1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc; free -g | head -2; lscpu | grep -E 'Model name|Socket|NUMA|^CPU\(s\)'
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log; lap pytest
CP360_AUTOTUNE_VERBOSE=1 timeout 400 python tools/site_sweep.py --cube 224 --batch 32,8,1 > $OUT/sweep224.txt 2> $OUT/sweep224.err; echo "sweep224 rc=$?"; cat $OUT/sweep224.txt; lap sweep224
timeout 300 python tools/site_sweep.py --cube 256 --batch 8,1 > $OUT/sweep256.txt 2> $OUT/sweep256.err; echo "sweep256 rc=$?"; cat $OUT/sweep256.txt; lap sweep256
for b in 1 2 4 8 16; do
  timeout 200 python bench.py --batch $b --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_b$b.json 2> $OUT/bench_b$b.err; echo "bench b$b rc=$?"; cut -c1-400 $OUT/bench_b$b.json
done; lap batchcurve
export CP360_AUTOTUNE=0
SMALL='not full_size and not resnet50_sites and not selftest and not model'
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $OUT/memcheck_pytest.log; grep -c "ERROR SUMMARY: 0 errors" $OUT/memcheck.log; tail -2 $OUT/memcheck.log; lap memcheck
for tool in racecheck synccheck; do
  timeout 360 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "($SMALL) and (cubepad_vs_oracle or c2e_vs_oracle or backward_cube_tile or cubic_vs_oracle)" \
      > $OUT/${tool}_pytest.log 2>&1; echo "$tool rc=$?"
  tail -3 $OUT/${tool}_pytest.log; tail -2 $OUT/$tool.log; lap $tool
done
ls -la $OUT
