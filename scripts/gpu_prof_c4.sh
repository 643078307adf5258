#!/bin/bash
# ncu --set full capture of the two-rows-per-lane solve kernel on the coupled gas + surface ensemble.  Usage: bash scripts/gpu_prof_c4.sh tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_c4.py > $OUT/plain_c4_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o $OUT/prof_c4_$TAG python scripts/prof_c4.py > $OUT/ncu_c4_$TAG.log 2>&1
tail -2 $OUT/ncu_c4_$TAG.log; tail -1 $OUT/plain_c4_$TAG.log
