#!/bin/bash
# Round 2, call 2: a-7 (reference call sites on the GPU) + racecheck with ring wrap (product and diagnostic build) + memcheck.
TAG=${1:-r2c2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests/test_reference_callsites_gpu.py -m gpu -q -x > $OUT/pytest_callsites.log 2>&1; echo "callsites rc=$?"; tail -25 $OUT/pytest_callsites.log; lap callsites
timeout 120 python tools/racecheck_probe.py > $OUT/probe_plain.log 2>&1; echo "probe (no tool) rc=$?"; cat $OUT/probe_plain.log; lap probe
export CP360_AUTOTUNE=0
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe.log python tools/racecheck_probe.py > $OUT/racecheck_probe.out 2>&1; echo "racecheck probe rc=$?"
tail -3 $OUT/racecheck_probe.out; tail -2 $OUT/racecheck_probe.log; lap racecheck_probe
CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_arriveall.so timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_probe_arriveall.log python tools/racecheck_probe.py --only cube,bwd > $OUT/racecheck_probe_arriveall.out 2>&1; echo "racecheck probe (arrive-all build) rc=$?"
tail -3 $OUT/racecheck_probe_arriveall.out; tail -2 $OUT/racecheck_probe_arriveall.log; lap racecheck_arriveall
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size and not resnet50_sites and not selftest and not model and not autotuned" > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $OUT/memcheck_pytest.log; tail -2 $OUT/memcheck.log; lap memcheck
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_probe.log python tools/racecheck_probe.py > $OUT/memcheck_probe.out 2>&1; echo "memcheck probe rc=$?"
tail -2 $OUT/memcheck_probe.out; tail -2 $OUT/memcheck_probe.log; lap memcheck_probe
for f in $OUT/racecheck*.log; do echo "$f: $(grep -c 'Race reported' $f) race records"; done
ls -la $OUT
