#!/bin/bash
# final two-GPU check of the driver's launch lines (own arm, reference arm, corpus workload)
TAG=${1:-r2n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/bench_chain_n2.json 2> $OUT/bench_chain_n2.err; echo "chain n=2 rc=$?"; cut -c1-400 $OUT/bench_chain_n2.json
timeout 300 $TR --nproc-per-node 2 --master-port 29703 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_reference_n2.json 2> $OUT/bench_reference_n2.err; echo "reference n=2 rc=$?"; cut -c1-300 $OUT/bench_reference_n2.json
timeout 300 $TR --nproc-per-node 2 --master-port 29704 bench.py --gpus 2 --workload corpus > $OUT/bench_corpus_n2.json 2> $OUT/bench_corpus_n2.err; echo "corpus n=2 rc=$?"; cut -c1-300 $OUT/bench_corpus_n2.json
timeout 300 python -m pytest tests -m gpu -q -x -k "shard or multi or rank or distributed" 2>&1 | tail -2
