#!/bin/bash
# synccheck on the c2e kernels after the predicated mbarrier init in the cluster kernel; then the rest of the probe
TAG=${1:-r2c36}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_c2e.log python tools/racecheck_probe.py --only c2e > $OUT/synccheck_c2e.out 2>&1; echo "synccheck c2e rc=$?"; tail -2 $OUT/synccheck_c2e.log; grep -c Divergent $OUT/synccheck_c2e.log
CP360_C2E_CLUSTER_SIZE=2 timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_all_cl2.log python tools/racecheck_probe.py > $OUT/synccheck_all_cl2.out 2>&1; echo "synccheck whole probe (cluster size 2) rc=$?"; tail -1 $OUT/synccheck_all_cl2.log; tail -3 $OUT/synccheck_all_cl2.out
