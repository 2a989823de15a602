#!/bin/bash
OUT=gpurun_out/${1:-diag_row3}; mkdir -p $OUT
export CP360_PDL=0
sw() { echo "== $*"; env "$@" timeout 300 python tools/kbench.py --only cubepad --iters 30 2>&1 | grep -E "row"; }
{
for rb in 10 14 18 19 21 22 23 26 28; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x64,128x64"; done
for rb in 7 11; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x128"; done
for rb in 8 10 12 14 16 18 20 24 28; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x56,128x56"; done
for rb in 7 8 9 10 11 12 14; do sw CP360_ROW_RB=$rb CP360_KB_SITES="64x112"; done
for rb in 3 4 5 6 7; do sw CP360_ROW_RB=$rb CP360_KB_SITES="3x224"; done
for rb in 8 12 16 20 24; do sw CP360_ROW_RB=$rb CP360_KB_SITES="128x32,256x32"; done
for rb in 7 10 14 20; do sw CP360_ROW_RB=$rb CP360_KB_SITES="128x28,256x28"; done
sw CP360_ROW_TILE_KB=8 CP360_KB_SITES="128x32,256x32,128x28,256x28"
sw CP360_ROW_TILE_KB=2 CP360_KB_SITES="128x32,256x32,128x28,256x28"
} 2>&1 | tee $OUT/sweep.txt
