#!/bin/bash
# ncu --set full capture of the forward-sensitivity kernel on the C5 ensemble.  Usage: bash scripts/gpu_prof_sens.sh tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_sens.py > $OUT/plain_sens_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sens_kernel -s 1 -c 1 -f -o $OUT/prof_sens_$TAG python scripts/prof_sens.py > $OUT/ncu_sens_$TAG.log 2>&1
tail -2 $OUT/ncu_sens_$TAG.log; tail -1 $OUT/plain_sens_$TAG.log
