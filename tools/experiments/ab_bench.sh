#!/bin/bash
# A/B the chain bench under environment variants (GPU box). Each line of the config list is a set of env assignments.
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT; shift
run() { echo "== $*"; env "$@" CP360_BENCH_SITES=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 100 2>$OUT/err.txt | tee $OUT/last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f frames/s  ms/step %.4f  chain %.1f GB/s (%.3f)' % (d['value'], d['ms_per_step'], d['roofline']['chain_gbs'], d['roofline']['chain_frac']))
for k,v in d['kernels'].items(): print('   %-34s share %.3f  %7.1f GB/s  x%d  %.1f us' % (k, v['share'], v['gbs'], v['launches_per_step'], v['avg_us']))
"; grep -E 'site|autotune' $OUT/err.txt | sed 's/^/   /' | cut -c1-400; }
while IFS= read -r line; do [ -z "$line" ] && continue; run $line; done 2>&1 | tee $OUT/ab.txt
