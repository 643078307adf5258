#!/bin/bash
# C5 exploratory check.  Usage: bash scripts/gpu_sens.sh
mkdir -p gpurun_out
timeout 900 python scripts/gpu_sens_check.py 2>&1 | tee gpurun_out/sens_check.log | tail -60
