#!/bin/bash
# Round 2, second evidence pass (after the backward / bicubic kernel work): GPU suite, smoke, headline bench lines,
# reference arm, kbench of every kernel, ncu --set full of the section-8(f) rows, sanitizers.
# Usage (repo root on the box): bash tools/gpu_r2_profile_b.sh <tag>
TAG=${1:-r2profb}
OUT=gpurun_out/$TAG
mkdir -p $OUT/summary
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -2 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; lap smoke
CP360_BENCH_SITES=1 timeout 500 python bench.py > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; cut -c1-300 $OUT/bench_256.json; lap bench256
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $OUT/bench_reference.json; lap reference
for sec in cubepad bwd c2e e2c fused; do
  timeout 300 python tools/kbench.py --only $sec --batch 32 > $OUT/kbench_$sec.txt 2>&1; echo "kbench $sec rc=$?"
done; cat $OUT/kbench_*.txt > $OUT/kbench.txt; lap kbench
timeout 300 python tools/experiments/c2e_bwd_time.py > $OUT/c2e_bwd_time.txt 2>&1
timeout 300 python tools/experiments/cubepad_bwd_time.py > $OUT/cubepad_bwd_time.txt 2>&1; lap bwd_sweeps
CP360_PROF_CUBE=0 timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -f -o $OUT/full_0 \
    python tools/prof_all.py 32 $OUT/full_0.order > $OUT/full_0.log 2>&1; echo "ncu full (f rows) rc=$?"; tail -1 $OUT/full_0.log
timeout 200 ncu -i $OUT/full_0.ncu-rep --page raw --csv > $OUT/full_0.raw.csv 2>/dev/null
rm -f $OUT/full_0.ncu-rep; lap full
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe.log python tools/racecheck_probe.py > $OUT/racecheck_probe.out 2>&1; echo "racecheck probe rc=$?"
  CP360_BWD_TABLE_CACHE=0 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe_bwd_nocache.log python tools/racecheck_probe.py --only bwd > $OUT/racecheck_probe_bwd_nocache.out 2>&1; echo "racecheck probe (tables built per CTA) rc=$?" )
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_suite.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model and (cubepad_vs_oracle or c2e_vs_oracle or backward_cube_tile or cubic_vs_oracle or nan_semantics or fused or team_split or table_cache)" > $OUT/racecheck_suite.out 2>&1; echo "racecheck suite rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_suite.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model" > $OUT/memcheck_suite.out 2>&1; echo "memcheck suite rc=$?"
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_probe.log python tools/racecheck_probe.py > $OUT/synccheck_probe.out 2>&1; echo "synccheck probe rc=$?"
for f in $OUT/racecheck*.log $OUT/memcheck*.log $OUT/synccheck*.log; do echo "$f: $(tail -1 $f)"; done; lap sanitizers
