#!/bin/bash
# ncu --set full capture of solve_kernel<1> on the ConstantTP ensemble (C1).  Usage: bash scripts/gpu_prof_c1.sh tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_c1.py > $OUT/plain_c1_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o $OUT/prof_c1_$TAG python scripts/prof_c1.py > $OUT/ncu_c1_$TAG.log 2>&1
tail -2 $OUT/ncu_c1_$TAG.log; tail -1 $OUT/plain_c1_$TAG.log
