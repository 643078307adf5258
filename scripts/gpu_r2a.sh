#!/bin/bash
# round 2, first GPU pass of the group-per-reactor kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== small check"; timeout 600 python scripts/gpu_small_check.py 20000 2>&1 | tail -12
echo "== memcheck (n=48)"; timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_small_check.py 48 2>&1 | tail -8
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_r2a.json 2> $OUT/bench_r2a.err; tail -c 3000 $OUT/bench_r2a.json; tail -3 $OUT/bench_r2a.err
