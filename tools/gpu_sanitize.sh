#!/bin/bash
# compute-sanitizer over the GPU parity suite (GPU box; not yet run in round 1 — see profiles/README.md "Known headroom").
# memcheck on the whole -m gpu suite minus the full-size cases (the tool slows kernels ~10-50x), then racecheck and
# synccheck on the kernels with hand-rolled mbarrier / shared-memory pipelines (row, cube-tile, c2e small, bwd cube-tile).
# The first-call autotuner is switched off so that every test exercises one deterministic tiling.
# Usage (from the repo root on the box): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export CP360_AUTOTUNE=0
SMALL='not full_size and not resnet50_sites and not selftest and not model'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $OUT/memcheck_pytest.log; grep -c "ERROR SUMMARY: 0 errors" $OUT/memcheck.log
for tool in racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "($SMALL) and (cubepad_vs_oracle or c2e_vs_oracle or backward_cube_tile or cubic_vs_oracle)" \
      > $OUT/${tool}_pytest.log 2>&1; echo "$tool rc=$?"
  tail -3 $OUT/${tool}_pytest.log; tail -2 $OUT/$tool.log
done
