#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 8): upload ceiling per GPU count, chain and corpus lines at 8 / 4 / 2 GPUs.
TAG=${1:-r2multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -E 'Model name|Socket|NUMA|^CPU\(s\)' > $OUT/host.txt; free -g | head -2 >> $OUT/host.txt; cat $OUT/host.txt
for n in 8 4 2 1; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29600 + n)) tools/h2d_probe.py --seconds 1.0 > $OUT/h2d_probe_n$n.txt 2> $OUT/h2d_probe_n$n.err; echo "h2d probe n=$n rc=$?"; grep '"probe"' $OUT/h2d_probe_n$n.txt; lap h2d_$n
done
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29700 + n)) bench.py --gpus $n --steps 100 --no-cpu-baseline --no-aten-baseline > $OUT/bench_chain_n$n.json 2> $OUT/bench_chain_n$n.err; echo "chain n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_chain_n$n.json")); e=d["e2e"]; print("chain n=$n value", d["value"], "e2e", e["value"], "h2d agg", e["h2d_gbs_aggregate"], "ceiling agg", e["h2d_ceiling_gbs_aggregate"], "frac", e["frac_of_h2d_ceiling"])
except Exception as ex: print("no line", ex)
PY
  lap chain_$n
done
for alloc in hugepage write_combined; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29811 bench.py --gpus 8 --steps 50 --no-cpu-baseline --no-aten-baseline --no-fused --host-alloc $alloc > $OUT/bench_chain_n8_$alloc.json 2> $OUT/bench_chain_n8_$alloc.err; echo "chain n=8 $alloc rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_chain_n8_$alloc.json")); e=d["e2e"]; print("chain n=8 $alloc e2e", e["value"], "ceiling agg", e["h2d_ceiling_gbs_aggregate"])
except Exception as ex: print("no line", ex)
PY
done; lap alloc_variants
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29900 + n)) bench.py --gpus $n --workload corpus --no-cpu-baseline > $OUT/bench_corpus_n$n.json 2> $OUT/bench_corpus_n$n.err; echo "corpus n=$n rc=$?"; cut -c1-180 $OUT/bench_corpus_n$n.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_corpus_n$n.json")); print("corpus n=$n", d["value"], "frames/s", d["corpus_wall_ms"], "ms/pass, gather", d["gather_ms"], "ms")
except Exception as ex: print("no line", ex)
PY
  lap corpus_$n
done
timeout 200 $TR --nproc-per-node 8 --master-port 29950 bench.py --gpus 8 --workload clstm --no-cpu-baseline > $OUT/bench_clstm_n8.json 2> $OUT/bench_clstm_n8.err; echo "clstm n=8 rc=$?"; cut -c1-200 $OUT/bench_clstm_n8.json; lap clstm8
ls -la $OUT
