#!/bin/bash
# ncu full capture of the solve kernel on a 20k-reactor ensemble.  Usage: bash scripts/gpu_prof.sh tag
set -u
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --n-per-gpu 20000 --no-cpu"
timeout 600 $CMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 1 -f -o $OUT/prof_solve_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log; python - <<PY
import json
print(json.loads(open("$OUT/plain_$TAG.log").read().strip().splitlines()[-1])["value"])
PY
