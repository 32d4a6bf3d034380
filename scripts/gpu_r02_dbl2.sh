#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in emul0 emul1 emul2 emul3; do
  if [ $v = emul2 ]; then unset FA2_SM100_LIB; else export FA2_SM100_LIB=variants/lib_$v.so; fi
  echo "== $v: $(FA2_KT_MODES=3 timeout 300 python scripts/kernel_times.py 2 16 16384 64 0 10 2>&1 | tail -1)"
done
done 2>&1 | tee gpurun_out/dbl_emul.txt
