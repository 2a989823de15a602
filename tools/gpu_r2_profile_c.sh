#!/bin/bash
# Round 2, final evidence pass on one B200 (20-launch chain: first site fused into e2c; table re-ranked inside it):
# GPU suite, smoke, bench lines (256 / 224 / two-launch first site / reference arm), batch curve, ConvLSTM and corpus lines,
# ncu launch lists, memcheck + racecheck of the ring-wrap probe.
TAG=${1:-r2profc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -2 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; lap smoke
CP360_BENCH_SITES=1 timeout 500 python bench.py > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; cut -c1-200 $OUT/bench_256.json; lap bench256
CP360_BENCH_SITES=1 timeout 500 python bench.py --cube 224 > $OUT/bench_224.json 2> $OUT/bench_224.err; echo "bench 224 rc=$?"; cut -c1-200 $OUT/bench_224.json; lap bench224
CP360_BENCH_SITES=1 timeout 500 python bench.py --no-fuse-first-site --no-cpu-baseline --no-aten-baseline > $OUT/bench_256_two_launch_first_site.json 2> $OUT/bench_256_two.err; echo "bench 256 (21 launches) rc=$?"; cut -c1-200 $OUT/bench_256_two_launch_first_site.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cut -c1-200 $OUT/bench_reference.json; lap reference
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
timeout 300 python bench.py --workload clstm > $OUT/bench_clstm.json 2> $OUT/bench_clstm.err; echo "clstm rc=$?"; cut -c1-200 $OUT/bench_clstm.json
timeout 300 python bench.py --workload clstm --clstm-variant reference > $OUT/bench_clstm_reference_widths.json 2> /dev/null; echo "clstm(ref widths) rc=$?"; cut -c1-200 $OUT/bench_clstm_reference_widths.json
timeout 300 python bench.py --workload corpus > $OUT/bench_corpus_n1.json 2> $OUT/bench_corpus.err; echo "corpus rc=$?"; cut -c1-200 $OUT/bench_corpus_n1.json; lap workloads
for cube in 256 224; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches_$cube.csv \
      python bench.py --cube $cube --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-aten-baseline --no-fused --profile-range > $OUT/bench_under_ncu_$cube.log 2>&1; echo "ncu list $cube rc=$?"
done; lap launches
( export CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe.log python tools/racecheck_probe.py > $OUT/racecheck_probe.out 2>&1; echo "racecheck probe rc=$?" )
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_probe.log python tools/racecheck_probe.py > $OUT/memcheck_probe.out 2>&1; echo "memcheck probe rc=$?"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_pipeline.log python -m pytest tests/test_pipeline_gpu.py -m gpu -q > $OUT/memcheck_pipeline.out 2>&1; echo "memcheck pipeline tests rc=$?"
for f in $OUT/racecheck*.log $OUT/memcheck*.log; do echo "$f: $(tail -1 $f)"; done; lap sanitizers
