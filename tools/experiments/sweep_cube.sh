#!/bin/bash
# Sweep cube-tile kernel knobs on the small-plane sites (GPU box).
OUT=gpurun_out/${1:-sweep_cube}; mkdir -p $OUT
export CP360_KB_SITES="128x32,256x32,256x16,512x16,512x8,2048x8,256x14,512x7,2000x7"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cubepad" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for cfg in "24 4 2 8" "24 4 2 16" "16 4 2 8" "32 3 2 8" "48 3 1 16" "12 6 2 8" "24 4 2 4" "24 3 3 8" "16 4 3 8"; do
  set -- $cfg
  echo "== stage_kb=$1 stages=$2 ctas=$3 warps=$4"
  CP360_CUBE_STAGE_KB=$1 CP360_CUBE_STAGES=$2 CP360_CUBE_CTAS=$3 CP360_CUBE_WARPS=$4 timeout 300 python tools/kbench.py --only cubepad 2>&1 | grep -E "cube2|row \*|cube \*"
done 2>&1 | tee $OUT/sweep.txt
