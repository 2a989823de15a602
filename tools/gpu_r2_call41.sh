#!/bin/bash
# does a short run (--steps 20 --warmup 3, the driver's flags in round 1) read lower than a long one on the same box?
OUT=gpurun_out/${1:-r2c41}; mkdir -p $OUT
Q="--no-cpu-baseline --no-e2e --no-aten-baseline --no-fused"
for cfg in "20 3" "20 50" "200 10" "20 3" "200 10"; do set -- $cfg
  python bench.py --steps $1 --warmup $2 $Q > $OUT/b_$1_$2.json 2>/dev/null
  python -c "import json; d=json.load(open('$OUT/b_$1_$2.json')); print('steps $1 warmup $2: %.1f frames/s %.4f ms dominant %.3f chain %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['chain_frac'], d['clocks']))"
done
