#!/bin/bash
# large-path block-size variants: build/variants/lib_lt*.so.  Usage: bash scripts/gpu_large_variants.sh [n_reactors] [ns]
for f in build/variants/lib_lt*.so; do
  echo "== $f"
  RMSB200_LIB=$PWD/$f timeout 600 python scripts/gpu_large_check.py ${2:-1000} ${1:-592} 2>&1 | grep "profile\|ensemble\|tend\|status"
done
