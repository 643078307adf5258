#!/bin/bash
# Round-end pass: parity tests, smoke, bench (both arms), ncu launch list of the bench command.  Usage: bash scripts/gpu_final.sh tag
set -u
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; tail -c 1200 $OUT/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
echo "== ncu launch list"
timeout 600 $CMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
tail -2 $OUT/ncu_launch_$TAG.log; wc -l $OUT/launches_$TAG.csv
