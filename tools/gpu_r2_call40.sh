#!/bin/bash
# per-line profile of K3m (c2e + channel max, cluster kernel) at [192,1000,8,8]
TAG=${1:-r2c40}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'c2e_max' -f -o $OUT/k3m python tools/prof_one.py c2emax 8 1000 32 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/ncu_summary.py $OUT/k3m.ncu-rep 2>&1 | grep -E "time_duration|inst_executed.sum|issue_active|stall|bank" | head -12
timeout 100 python tools/ncu_lines.py $OUT/k3m.ncu-rep 40 > $OUT/k3m_lines.txt 2>&1; rm -f $OUT/k3m.ncu-rep; head -36 $OUT/k3m_lines.txt | cut -c1-170
