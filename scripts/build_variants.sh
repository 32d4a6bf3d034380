#!/bin/bash
# Builds tuning variants of libfa2_sm100.so into variants/ (git-ignored; they travel to the GPU box):
#   scripts/build_variants.sh name "file1.cu file2.cu" "-DFLAG=.." [name files flags ...]
# Only the listed sources are recompiled with the extra flags; every other object is taken from the regular build
# (run `make -C fa2_triton_b200/csrc` first). Select a variant at run time with FA2_SM100_LIB=variants/lib_<name>.so.
set -e
cd "$(dirname "$0")/../fa2_triton_b200/csrc"
mkdir -p ../../variants
ALL="api fwd_sm100 fwd2_sm100 bwd_sm100 bwd_dkdv_sm100 bwd_dq_sm100 bwd_dq_spill_sm100 bwd256_sm100 varlen_sm100"
while [ $# -gt 2 ]; do
  name=$1; files=$2; flags=$3; shift 3
  d=../../variants/obj_$name; mkdir -p $d
  objs=""
  for f in $ALL; do  # (probe_sm100 is a test-only library of its own)
    if echo " $files " | grep -q " $f.cu "; then
      nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $f.cu -o $d/$f.o 2> $d/$f.log &
      objs="$objs $d/$f.o"
    else
      objs="$objs $f.o"
    fi
  done
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../variants/lib_$name.so $objs -lcudart
  rm -rf $d
  echo built variants/lib_$name.so
done
