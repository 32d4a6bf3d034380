#!/bin/bash
# Round 2: forward exponentials that do not wait for the block maximum (fwd2_sm100.cu) - parity, A/B timing, trace.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_varlen_headdim_gpu.py tests/test_triton_golden_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/fwd_tests.log 2>&1
echo "fwd tests exit $?"; tail -3 gpurun_out/fwd_tests.log
for rep in 1 2; do
  echo "== before"; FWD_TIMES_CASES=4 FA2_SM100_LIB=variants/lib_notoken.so timeout 300 python scripts/fwd_times.py 2>&1 | tee -a gpurun_out/fwd_times_before.txt
  echo "== speculative"; FWD_TIMES_CASES=4 timeout 300 python scripts/fwd_times.py 2>&1 | tee -a gpurun_out/fwd_times_spec.txt
done
timeout 300 python scripts/fwd2_trace.py 0 > gpurun_out/fwd2_trace_spec.txt 2>&1; sed -n 1,24p gpurun_out/fwd2_trace_spec.txt
