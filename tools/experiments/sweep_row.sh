#!/bin/bash
# Sweep row-kernel knobs on the mid-size sites (GPU box) + ncu full captures of the W=64 cases.
OUT=gpurun_out/${1:-sweep_row}; mkdir -p $OUT
export CP360_KB_SITES="64x64,128x64,64x128,3x256"
sw() { echo "== $*"; env "$@" timeout 300 python tools/kbench.py --only cubepad --iters 30 2>&1 | grep -E "row"; }
{
sw CP360_PDL=0
sw CP360_PDL=0 CP360_ROW_BALANCE=0
for rb in 8 12 16 20 24 32; do sw CP360_PDL=0 CP360_ROW_RB=$rb CP360_KB_SITES="64x64,128x64"; done
for rb in 4 6 8 9 10 12 16; do sw CP360_PDL=0 CP360_ROW_RB=$rb CP360_KB_SITES="64x128"; done
for rb in 2 3 4 5 6 8; do sw CP360_PDL=0 CP360_ROW_RB=$rb CP360_KB_SITES="3x256"; done
sw CP360_PDL=0 CP360_ROW_UNIT_BANDS=2
sw CP360_PDL=0 CP360_ROW_UNIT_BANDS=4
sw CP360_PDL=0 CP360_ROW_SLOTS=4
sw CP360_PDL=0 CP360_ROW_SLOTS=2
sw CP360_PDL=0 CP360_ROW_CTAS=2
sw CP360_PDL=0 CP360_ROW_CTAS=2 CP360_ROW_SLOTS=2
sw CP360_PDL=0 CP360_ROW_TILE_KB=6
sw CP360_PDL=0 CP360_ROW_TILE_KB=8 CP360_ROW_SLOTS=2
} 2>&1 | tee $OUT/sweep.txt
for spec in "cubepad 64 64 1" "cubepad 128 64 1"; do
  name=$(echo $spec | tr ' ' '_')
  CP360_PDL=0 timeout 600 ncu --set full --clock-control none --import-source on -c 1 -s 2 -k regex:'cubepad' -f -o $OUT/full_$name \
      python tools/prof_one.py $spec > $OUT/full_$name.log 2>&1; echo "ncu full $name rc=$?"
done
