#!/bin/bash
TAG=${1:-r2c13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "c2e or cubic or pipeline or smoke" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_probe.log python tools/racecheck_probe.py > $OUT/synccheck_probe.out 2>&1; echo "synccheck probe rc=$?"; tail -1 $OUT/synccheck_probe.log; grep "Barrier error" -A 2 $OUT/synccheck_probe.log | head -6
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_c2e.log python tools/racecheck_probe.py --only c2e > $OUT/racecheck_c2e.out 2>&1; echo "racecheck c2e rc=$?"; tail -1 $OUT/racecheck_c2e.log
