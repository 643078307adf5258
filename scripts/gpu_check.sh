#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list + one full capture of the solve kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 6000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; tail -c 1500 $OUT/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --n-per-gpu 20000 --no-cpu"
echo "== ncu"
timeout 600 $CMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 1 -f -o $OUT/prof_solve_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT
