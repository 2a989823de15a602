#!/bin/bash
# the chain with its first site fused into e2c (bench default) next to the two-launch form; pipeline tests
TAG=${1:-r2c37}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q > $OUT/pytest_pipeline.log 2>&1; echo "pytest pipeline rc=$?"; tail -3 $OUT/pytest_pipeline.log
show() { python - <<PY
import json
d=json.load(open("$OUT/$1.json")); e=d.get("e2e") or {}; f=d.get("fused_chain") or {}
print("$1: %.1f frames/s  %.4f ms  launches/step %d  dominant %s %.3f chain %.3f  bytes/frame %.2f MB | e2e %s | fused %s" % (d["value"], d["ms_per_step"], d["gpu_launches"] // d["steps"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["chain_frac"], d["config"]["algorithmic_bytes_per_frame"] / 1e6, e.get("value"), f.get("value")))
PY
}
CP360_BENCH_SITES=1 timeout 500 python bench.py > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "rc=$?"; show bench_256; grep -E "site (e2c|cubepad_row 3x)" $OUT/bench_256.err
CP360_BENCH_SITES=1 timeout 500 python bench.py --no-fuse-first-site --no-cpu-baseline --no-aten-baseline > $OUT/bench_256_two_launch_first_site.json 2> $OUT/bench_256_two.err; show bench_256_two_launch_first_site; grep -E "site (e2c|cubepad_row 3x)" $OUT/bench_256_two.err
CP360_BENCH_SITES=1 timeout 500 python bench.py --cube 224 --no-cpu-baseline --no-aten-baseline > $OUT/bench_224.json 2> $OUT/bench_224.err; show bench_224
timeout 300 python bench.py --workload corpus --no-cpu-baseline > $OUT/bench_corpus_n1.json 2> $OUT/bench_corpus.err; echo "corpus rc=$?"; cut -c1-200 $OUT/bench_corpus_n1.json
for b in 1 8; do timeout 200 python bench.py --batch $b --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-aten-baseline --no-fused > $OUT/bench_256_b$b.json 2>/dev/null; show bench_256_b$b; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches_256.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-aten-baseline --no-fused --profile-range > $OUT/bench_under_ncu_256.log 2>&1; echo "ncu list rc=$?"; grep -c "cp360" $OUT/launches_256.csv
