#!/bin/bash
# Round 2: the new forward kernel (fwd2_sm100.cu) - parity first, then timing at configs 2 / 4 and head dim 256.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_varlen_headdim_gpu.py tests/test_triton_golden_gpu.py tests/test_backward_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/fwd_tests.log 2>&1
echo "fwd tests exit $?"; tail -15 gpurun_out/fwd_tests.log
timeout 300 python scripts/fwd_times.py > gpurun_out/fwd_times.txt 2>&1; cat gpurun_out/fwd_times.txt
