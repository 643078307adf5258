#!/bin/bash
# full GPU test suite (no -x) + short bench.  Usage: bash scripts/gpu_tests.sh tag [pytest -k expression]
set -u
TAG=${1:-x}; KEXPR=${2:-}; OUT=gpurun_out; mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 2400 python -m pytest tests -m gpu -q -k "$KEXPR" 2>&1 | tail -60 > $OUT/pytest_$TAG.log
else
  timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_$TAG.log
fi
tail -5 $OUT/pytest_$TAG.log
timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
