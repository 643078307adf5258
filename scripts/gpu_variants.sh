#!/bin/bash
# time solve-kernel build variants: build/variants/lib_*.so, each through bench.py
for f in build/variants/lib_*.so; do
  echo -n "$f: "
  RMSB200_LIB=$PWD/$f timeout 300 python bench.py --steps 3 --warmup 3 --n-per-gpu 100000 --no-cpu --no-others 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'solves/s  e2e', round(d['e2e']['value']), {k: round(v, 1) for k, v in d['config']['mean_per_solve'].items()})"
done
