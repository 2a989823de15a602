#!/bin/bash
# Round 2, call 3: racecheck with deep ring wrap — product build vs the -DCP360_ARRIVE_ALL diagnostic build.
TAG=${1:-r2c3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
export CP360_AUTOTUNE=0 CP360_CUBE_STAGES=2 CP360_CUBE_STAGE_KB=24 CP360_BWD_STAGES=2 CP360_BWD_STAGE_KB=32
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_wrap.log python tools/racecheck_probe.py --only row,cube,bwd > $OUT/racecheck_wrap.out 2>&1; echo "racecheck (product build, 2-stage rings) rc=$?"
tail -2 $OUT/racecheck_wrap.out; tail -2 $OUT/racecheck_wrap.log; lap product
CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_arriveall.so timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_wrap_arriveall.log python tools/racecheck_probe.py --only cube,bwd > $OUT/racecheck_wrap_arriveall.out 2>&1; echo "racecheck (arrive-all build, 2-stage rings) rc=$?"
tail -2 $OUT/racecheck_wrap_arriveall.out; tail -2 $OUT/racecheck_wrap_arriveall.log; lap arriveall
for f in $OUT/racecheck*.log; do echo "$f: $(grep -c 'Race reported' $f) race records"; done
