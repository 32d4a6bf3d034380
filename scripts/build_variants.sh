#!/bin/bash
# Builds tuning variants of libfa2_sm100.so into variants/ (git-ignored): usage scripts/build_variants.sh name "-DFLAG=.." ...
set -e
cd "$(dirname "$0")/../fa2_triton_b200/csrc"
mkdir -p ../../variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  d=../../variants/obj_$name; mkdir -p $d
  for f in api fwd_sm100 bwd_sm100 bwd_dkdv_sm100 bwd_dq_sm100 bwd_dq_spill_sm100 varlen_sm100 probe_sm100; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $f.cu -o $d/$f.o &
  done
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../variants/lib_$name.so $d/*.o -lcudart
  rm -rf $d
  echo built variants/lib_$name.so
done
