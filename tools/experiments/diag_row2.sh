#!/bin/bash
OUT=gpurun_out/${1:-diag_row2}; mkdir -p $OUT
export CP360_PDL=0
sw() { echo "== $*"; env "$@" timeout 300 python tools/kbench.py --only cubepad --iters 30 2>&1 | grep -E "row"; }
{
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k cubepad 2>&1 | tail -1
CP360_ROW_ORDER=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k cubepad 2>&1 | tail -1
for rb in 8 12 16 20 24 32; do sw CP360_ROW_ORDER=1 CP360_ROW_RB=$rb CP360_KB_SITES="64x64,128x64"; done
for rb in 6 8 9 10 12 16; do sw CP360_ROW_ORDER=1 CP360_ROW_RB=$rb CP360_KB_SITES="64x128"; done
for rb in 3 4 5 6 8; do sw CP360_ROW_ORDER=1 CP360_ROW_RB=$rb CP360_KB_SITES="3x256"; done
sw CP360_ROW_ORDER=1 CP360_KB_SITES="128x32,256x32"
sw CP360_ROW_ORDER=0 CP360_KB_SITES="128x32,256x32"
sw CP360_ROW_ORDER=1 CP360_ROW_SLOTS=4 CP360_KB_SITES="64x64,128x64,64x128"
sw CP360_ROW_ORDER=1 CP360_ROW_SLOTS=2 CP360_KB_SITES="64x64,128x64,64x128"
} 2>&1 | tee $OUT/sweep.txt
