#!/bin/bash
# eight-GPU lines of the 20-launch chain (first site fused into e2c) and of the corpus pass
TAG=${1:-r2n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --steps 100 --warmup 5 --no-fused > $OUT/bench_chain_n8.json 2> $OUT/bench_chain_n8.err; echo "chain n=8 rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_chain_n8.json")); e=d["e2e"]; print("chain n=8 value", d["value"], "e2e", e["value"], "h2d agg", e["h2d_gbs_aggregate"], "ceiling agg", e["h2d_ceiling_gbs_aggregate"], "frac", e["frac_of_h2d_ceiling"])
PY
timeout 300 $TR --nproc-per-node 8 --master-port 29709 bench.py --gpus 8 --workload corpus > $OUT/bench_corpus_n8.json 2> $OUT/bench_corpus_n8.err; echo "corpus n=8 rc=$?"; cut -c1-260 $OUT/bench_corpus_n8.json
