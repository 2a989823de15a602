#!/bin/bash
OUT=gpurun_out/${1:-diag_row}; mkdir -p $OUT
export CP360_KB_SITES="64x64,128x64,64x128" CP360_PDL=0
sw() { echo "== $*"; env "$@" timeout 300 python tools/kbench.py --only cubepad --iters 30 2>&1 | grep -E "row"; }
{
for rb in 16 20; do for sl in 2 3 4 6; do sw CP360_ROW_RB=$rb CP360_ROW_SLOTS=$sl CP360_KB_SITES="64x64,128x64"; done; done
for st in 200 1000 4000; do sw CP360_ROW_RB=16 CP360_ROW_STAGGER_NS=$st CP360_KB_SITES="64x64,128x64"; done
for st in 1000; do sw CP360_ROW_RB=20 CP360_ROW_STAGGER_NS=$st CP360_KB_SITES="64x64,128x64"; done
for rb in 8 10; do for sl in 2 3 4; do sw CP360_ROW_RB=$rb CP360_ROW_SLOTS=$sl CP360_KB_SITES="64x128"; done; done
sw CP360_ROW_RB=10 CP360_ROW_STAGGER_NS=1000 CP360_KB_SITES="64x128"
} 2>&1 | tee $OUT/sweep.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_requests_srcunit_tex_op_write.sum
for rb in 16 20; do
  CP360_PROF_FLUSH=1 CP360_ROW_RB=$rb timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:cubepad_row -s 3 -c 2 --csv --log-file $OUT/regime_128_64_rb$rb.csv python tools/prof_one.py cubepad 128 64 1 > $OUT/regime_rb$rb.log 2>&1; echo "ncu rb=$rb rc=$?"
done
for rb in 8 10; do
  CP360_PROF_FLUSH=1 CP360_ROW_RB=$rb timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:cubepad_row -s 3 -c 2 --csv --log-file $OUT/regime_64_128_rb$rb.csv python tools/prof_one.py cubepad 64 128 1 > $OUT/regime_w128_rb$rb.log 2>&1; echo "ncu rb=$rb rc=$?"
done
