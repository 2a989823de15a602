#!/bin/bash
# Profile-only GPU-box pass (what tools/gpu_final.sh collects, sized to fit gpurun's 64 MiB return limit): ncu launch
# list of the bench command, ONE ncu --set full report over every kernel/shape of the chain (tools/prof_all.py), one
# more capture of the dominant row-kernel site with source import, summaries extracted on the box, per-kernel bench.
# Usage (from the repo root on the box): bash tools/gpu_profile.sh <tag> [frames per step]
TAG=${1:-prof}
B=${2:-32}
OUT=gpurun_out/$TAG
mkdir -p $OUT/summary
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --batch $B --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --profile-range > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; lap launches
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -f -o $OUT/full_all \
    python tools/prof_all.py $B $OUT/full_all.order > $OUT/full_all.log 2>&1; echo "ncu full (all sites) rc=$?"; tail -1 $OUT/full_all.log; lap full_all
CP360_PROF_ONLY=cubepad_64_128_1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubepad' -f \
    -o $OUT/full_cubepad_64_128_1 python tools/prof_all.py $B > $OUT/full_cubepad_64_128_1.log 2>&1; echo "ncu full + source (row kernel 64x128) rc=$?"; lap full_src
for r in $OUT/*.ncu-rep; do timeout 120 ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done; lap raw_csv
timeout 120 python tools/make_profiles.py $OUT $TAG $B $OUT/summary > $OUT/summary/make_profiles.log 2>&1; echo "summaries rc=$?"; lap summaries
timeout 120 ncu -i $OUT/full_cubepad_64_128_1.ncu-rep --page source --csv > $OUT/summary/row_kernel_64_128_source.csv 2>/dev/null; lap source_page
for sec in bwd c2e e2c fused; do
  timeout 200 python tools/kbench.py --only $sec > $OUT/kbench_$sec.txt 2>&1; echo "kbench $sec rc=$?"; lap kbench_$sec
done
cat $OUT/kbench_*.txt > $OUT/kbench.txt
# stay under the return limit: drop the big reports first (their summaries are in summary/)
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do
  big=$(ls -S $OUT/*.ncu-rep 2>/dev/null | head -1)
  [ -z "$big" ] && break
  echo "dropping $big ($(du -m $big | cut -f1) MB) to fit the return limit"; rm -f $big
done
ls -la $OUT $OUT/summary; du -sm gpurun_out
