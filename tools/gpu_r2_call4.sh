#!/bin/bash
# Round 2, call 4: racecheck arrive-all experiment, full GPU suite, and the new bench workloads.
TAG=${1:-r2c4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log; lap pytest
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log; lap smoke
( export CP360_AUTOTUNE=0 CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
  CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_arriveall.so timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_wrap_arriveall.log python tools/racecheck_probe.py --only cube,bwd > $OUT/racecheck_wrap_arriveall.out 2>&1; echo "racecheck (arrive-all build, 2-stage rings) rc=$?"
  tail -2 $OUT/racecheck_wrap_arriveall.out; tail -2 $OUT/racecheck_wrap_arriveall.log
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_row_c2e.log python tools/racecheck_probe.py --only row,c2e,e2c > $OUT/racecheck_row_c2e.out 2>&1; echo "racecheck (product build: row, c2e, e2c) rc=$?"
  tail -2 $OUT/racecheck_row_c2e.out; tail -2 $OUT/racecheck_row_c2e.log ); lap racecheck
CP360_BENCH_SITES=1 timeout 400 python bench.py --steps 100 > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?"; cut -c1-1500 $OUT/bench_256.json; tail -30 $OUT/bench_256.err; lap bench256
CP360_BENCH_SITES=1 timeout 400 python bench.py --cube 224 --steps 100 --no-cpu-baseline > $OUT/bench_224.json 2> $OUT/bench_224.err; echo "bench 224 rc=$?"; cut -c1-600 $OUT/bench_224.json; tail -30 $OUT/bench_224.err; lap bench224
timeout 300 python bench.py --workload clstm > $OUT/bench_clstm.json 2> $OUT/bench_clstm.err; echo "bench clstm rc=$?"; cat $OUT/bench_clstm.json; tail -5 $OUT/bench_clstm.err; lap clstm
timeout 300 python bench.py --workload clstm --clstm-variant reference > $OUT/bench_clstm_ref.json 2> $OUT/bench_clstm_ref.err; echo "bench clstm(reference widths) rc=$?"; cat $OUT/bench_clstm_ref.json; tail -5 $OUT/bench_clstm_ref.err; lap clstm_ref
timeout 300 python bench.py --workload corpus --no-cpu-baseline > $OUT/bench_corpus.json 2> $OUT/bench_corpus.err; echo "bench corpus rc=$?"; cat $OUT/bench_corpus.json; tail -5 $OUT/bench_corpus.err; lap corpus
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cat $OUT/bench_reference.json; lap reference
timeout 100 python tools/h2d_probe.py > $OUT/h2d_probe_n1.txt 2>&1; echo "h2d probe rc=$?"; cat $OUT/h2d_probe_n1.txt; lap h2d
ls -la $OUT
