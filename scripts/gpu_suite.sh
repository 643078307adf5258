#!/bin/bash
# tests + smoke + 1-GPU bench (both arms).  Usage: bash scripts/gpu_suite.sh tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3000 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
