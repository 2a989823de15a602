#!/bin/bash
TAG=${1:-r2c11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python tools/kbench.py --only c2e --batch 32 > $OUT/kbench_c2e.txt 2>&1; echo "kbench c2e rc=$?"; grep -E "cubic" $OUT/kbench_c2e.txt
CP360_CUBIC_T=0 timeout 300 python tools/kbench.py --only c2e --batch 32 2>&1 | grep -E "cubic"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_cubic.log python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cubic" > $OUT/memcheck_cubic.out 2>&1; echo "memcheck cubic rc=$?"; tail -1 $OUT/memcheck_cubic.log
