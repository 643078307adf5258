#!/bin/bash
# build/variants/lib_<name>.so: rmsb_small.cu compiled with the given -D flags, linked against the current rmsb_api.o.
# Usage: bash scripts/build_small_variants.sh name "-DRMSB_SMALL_TICKSYNC=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build_cuda()"
mkdir -p build/variants
C=reactionmechanismsimulator.jl_b200/csrc
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c -o build/variants/small_$name.o $C/rmsb_small.cu &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so build/rmsb_api.o build/rmsb_spec.o build/variants/small_$name.o && echo built $name ) &
done
wait
