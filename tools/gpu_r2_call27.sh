#!/bin/bash
TAG=${1:-r2c27}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/experiments/cubepad_bwd_time.py | tee $OUT/bwd_B.txt
