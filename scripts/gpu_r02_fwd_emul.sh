#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_varlen_headdim_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/fwd_tests.log 2>&1
echo "fwd tests exit $?"; tail -3 gpurun_out/fwd_tests.log
for v in "" emul1 emul2; do
  echo "== variant ${v:-default}"
  if [ -n "$v" ]; then export FA2_SM100_LIB=variants/lib_$v.so; else unset FA2_SM100_LIB; fi
  FWD_TIMES_CASES=4 timeout 300 python scripts/fwd_times.py 2>&1 | tee -a gpurun_out/fwd_times_emul.txt
done
