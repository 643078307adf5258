#!/bin/bash
# full GPU test suite (no -x) + short bench.  Usage: bash scripts/gpu_tests.sh tag
set -u
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/pytest_$TAG.log; tail -5 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
