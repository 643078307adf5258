#!/bin/bash
# ncu --set full capture of the solve kernel on the bench workload (100k reactors).  Usage: bash scripts/gpu_prof_full.sh tag
set -u
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-others"
timeout 600 $CMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 1 -f -o $OUT/prof_solve_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log; tail -c 600 $OUT/plain_$TAG.log
