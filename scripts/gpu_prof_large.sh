#!/bin/bash
# ncu --set full capture of the block-per-reactor solve kernel.  Usage: bash scripts/gpu_prof_large.sh tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_large.py > $OUT/plain_large_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:large_solve_kernel -s 1 -c 1 -f -o $OUT/prof_large_$TAG python scripts/prof_large.py > $OUT/ncu_large_$TAG.log 2>&1
tail -2 $OUT/ncu_large_$TAG.log; tail -1 $OUT/plain_large_$TAG.log
