#!/bin/bash
for fl in 3 1 2 0; do
  echo -n "flags=$fl: "
  timeout 300 python bench.py --steps 3 --warmup 3 --n-per-gpu 50000 --no-cpu --flags $fl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'solves/s', d['config']['mean_per_solve'], d['config']['all_converged'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
