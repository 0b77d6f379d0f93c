#!/bin/bash
# Developer tool: build libvc2b200 variants of idwt.cu (DD(9,7)-only dev build) for A/B timing:
#   tools/build_variant.sh NAME -DVC2B_IDWT16_MINB=20 ...   ->  vc2_conformance_b200/_lib/variant_NAME.so
set -e
cd "$(dirname "$0")/../vc2_conformance_b200"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -DVC2B_DEV_BUILD "$@" -Xptxas -v -c csrc/idwt.cu -o _lib/variant_$name.o 2> _lib/variant_$name.ptxas
grep -A3 "stream8_kernelILi0ELi0ELi64ELi3" _lib/variant_$name.ptxas | grep -E "Used|spill" | head -2
others=$(ls _lib/*.o | grep -v -e '_lib/idwt.o' -e '_lib/variant_')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _lib/variant_$name.so _lib/variant_$name.o $others
