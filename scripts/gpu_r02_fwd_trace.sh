#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/fwd_tests.log 2>&1
echo "fwd tests exit $?"; tail -3 gpurun_out/fwd_tests.log
FA2_SM100_LIB=fa2_triton_b200/libfa2_sm100_trace.so timeout 300 python scripts/fwd2_trace.py 0 > gpurun_out/fwd2_trace_noncausal.txt 2>&1
head -45 gpurun_out/fwd2_trace_noncausal.txt
timeout 300 python scripts/fwd_times.py > gpurun_out/fwd_times.txt 2>&1; head -3 gpurun_out/fwd_times.txt
