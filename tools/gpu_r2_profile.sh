#!/bin/bash
# Round 2 evidence pass (one B200): bench lines for every workload / face width / batch size, ncu launch lists,
# ncu --set full captures of every kernel/shape (256 family, 224 family, section-8(f) rows), sanitizer logs.
# Usage (repo root on the box): bash tools/gpu_r2_profile.sh <tag>
TAG=${1:-r2prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT/summary
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -2 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; lap smoke
CP360_BENCH_SITES=1 timeout 500 python bench.py > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; cut -c1-300 $OUT/bench_256.json; lap bench256
CP360_BENCH_SITES=1 timeout 500 python bench.py --cube 224 > $OUT/bench_224.json 2> $OUT/bench_224.err; echo "bench 224 rc=$?"; cut -c1-300 $OUT/bench_224.json; lap bench224
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; lap reference
for cube in 224 256; do for b in 1 2 4 8 16 32 64; do
  timeout 200 python bench.py --cube $cube --batch $b --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-aten-baseline > $OUT/bench_${cube}_b$b.json 2> /dev/null
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${cube}_b$b.json")); f=d.get("fused_chain") or {}
    print("cube $cube B %2d: %9.1f frames/s %8.4f ms/step chain %6.1f GB/s (%.3f) dominant %s %.3f fused %s frames/s" % ($b, d["value"], d["ms_per_step"], d["roofline"]["chain_gbs"], d["roofline"]["chain_frac"], d["roofline"]["kernel"], d["roofline"]["frac"], f.get("value")))
except Exception as e: print("cube $cube B $b: no line", e)
PY
done; done | tee $OUT/batch_curve.txt; lap batch_curve
timeout 300 python bench.py --workload clstm > $OUT/bench_clstm.json 2> $OUT/bench_clstm.err; echo "clstm rc=$?"; cut -c1-250 $OUT/bench_clstm.json
timeout 300 python bench.py --workload clstm --clstm-variant reference > $OUT/bench_clstm_reference_widths.json 2> /dev/null; echo "clstm(ref widths) rc=$?"; cut -c1-250 $OUT/bench_clstm_reference_widths.json
for b in 1 4; do timeout 200 python bench.py --workload clstm --clstm-variant reference --batch $b --no-cpu-baseline --no-e2e > $OUT/bench_clstm_reference_widths_b$b.json 2> /dev/null; cut -c1-200 $OUT/bench_clstm_reference_widths_b$b.json; done
timeout 300 python bench.py --workload corpus > $OUT/bench_corpus_n1.json 2> $OUT/bench_corpus.err; echo "corpus rc=$?"; cut -c1-250 $OUT/bench_corpus_n1.json; lap workloads
for cube in 256 224; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches_$cube.csv \
      python bench.py --cube $cube --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-aten-baseline --no-fused --profile-range > $OUT/bench_under_ncu_$cube.log 2>&1; echo "ncu list $cube rc=$?"
done; lap launches
for cube in 256 224 0; do
  CP360_PROF_CUBE=$cube timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -f -o $OUT/full_$cube \
      python tools/prof_all.py 32 $OUT/full_$cube.order > $OUT/full_$cube.log 2>&1; echo "ncu full (set $cube) rc=$?"; tail -1 $OUT/full_$cube.log
  timeout 200 ncu -i $OUT/full_$cube.ncu-rep --page raw --csv > $OUT/full_$cube.raw.csv 2>/dev/null
  rm -f $OUT/full_$cube.ncu-rep
done; lap full
CP360_PROF_CUBE=256 CP360_PROF_ONLY=cubepad_64_128_1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad' -f \
    -o $OUT/src_cubepad_64_128_1 python tools/prof_all.py 32 > $OUT/src_cubepad_64_128_1.log 2>&1; echo "ncu + source rc=$?"
timeout 120 python tools/ncu_lines.py $OUT/src_cubepad_64_128_1.ncu-rep 40 > $OUT/summary/row_kernel_64_128_lines.txt 2>&1; rm -f $OUT/src_cubepad_64_128_1.ncu-rep; lap source
for sec in cubepad bwd c2e e2c fused; do
  timeout 300 python tools/kbench.py --only $sec --batch 32 > $OUT/kbench_$sec.txt 2>&1; echo "kbench $sec rc=$?"
done; cat $OUT/kbench_*.txt > $OUT/kbench.txt; lap kbench
timeout 300 python tools/site_sweep.py --cube 224,256 --batch 32,8,1 > $OUT/site_sweep.txt 2>&1; echo "site sweep rc=$?"; lap sweep
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe.log python tools/racecheck_probe.py > $OUT/racecheck_probe.out 2>&1; echo "racecheck probe rc=$?"
  CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_perwarp.so timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe_perwarp.log python tools/racecheck_probe.py --only cube,bwd > $OUT/racecheck_probe_perwarp.out 2>&1; echo "racecheck probe (per-warp-arrive comparison build) rc=$?" )
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_suite.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model and (cubepad_vs_oracle or c2e_vs_oracle or backward_cube_tile or cubic_vs_oracle or nan_semantics or fused)" > $OUT/racecheck_suite.out 2>&1; echo "racecheck suite rc=$?"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_suite.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model" > $OUT/memcheck_suite.out 2>&1; echo "memcheck suite rc=$?"
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_probe.log python tools/racecheck_probe.py > $OUT/synccheck_probe.out 2>&1; echo "synccheck probe rc=$?"
for f in $OUT/racecheck*.log $OUT/memcheck*.log $OUT/synccheck*.log; do echo "$f: $(tail -1 $f)"; done; lap sanitizers
ls -la $OUT; du -sm gpurun_out
