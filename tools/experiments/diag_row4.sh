#!/bin/bash
OUT=gpurun_out/${1:-diag_row4}; mkdir -p $OUT
export CP360_PDL=0
sw() { echo "== $*"; env "$@" timeout 300 python tools/kbench.py --only cubepad --iters 30 2>&1 | grep -E "row"; }
{
export CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_stdef.so
for rb in 12 14 16 20 24; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x64,128x64"; done
for rb in 8 9 10 12; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x128"; done
sw CP360_KB_SITES="3x256,128x32,256x32"
} 2>&1 | tee $OUT/sweep.txt
