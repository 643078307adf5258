#!/bin/bash
# ncu --set full captures of the thread-per-reactor dydt / Jacobian kernels.  Usage: bash scripts/gpu_prof_evals.sh tag
set -u
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/prof_evals.py 3 > $OUT/plain_evals_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tpr_kernel|rmsb_spec" -s 2 -c 2 -f -o $OUT/prof_evals_$TAG python scripts/prof_evals.py 3 > $OUT/ncu_evals_$TAG.log 2>&1
tail -3 $OUT/plain_evals_$TAG.log; tail -2 $OUT/ncu_evals_$TAG.log
