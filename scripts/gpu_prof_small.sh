#!/bin/bash
# ncu --set full capture of the group-per-reactor solve kernel.  Usage: bash scripts/gpu_prof_small.sh tag [n]
set -u
TAG=${1:-x}; N=${2:-20000}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-others --n-per-gpu $N"
timeout 600 $CMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:small_solve_kernel -s 3 -c 1 -f -o $OUT/prof_small_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log; tail -c 400 $OUT/plain_$TAG.log
