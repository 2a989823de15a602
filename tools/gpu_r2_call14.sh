#!/bin/bash
TAG=${1:-r2c14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for g in 1 2 8; do
  CP360_C2E_CLUSTER_SIZE=$g timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_c2e_g$g.log python tools/racecheck_probe.py --only c2e > $OUT/synccheck_c2e_g$g.out 2>&1; echo "synccheck c2e cluster size $g rc=$?"; tail -1 $OUT/synccheck_c2e_g$g.log; grep "by thread" $OUT/synccheck_c2e_g$g.log | sed 's/.*in block/block/' | sort | uniq -c | sort -k2 | head -12
done
for g in 1 2 4 8; do echo "cluster size $g"; CP360_C2E_CLUSTER_SIZE=$g timeout 200 python tools/kbench.py --only c2e --batch 32 2>&1 | grep -E "c2e\+max  \[192,(1000|2048),[78]"; done
