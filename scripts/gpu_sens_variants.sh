#!/bin/bash
for f in build/variants/lib_sens*.so; do
  echo "== $f"
  RMSB200_LIB=$PWD/$f timeout 300 python scripts/gpu_sens_check.py 2>&1 | grep "C5 ensemble"
done
