#!/bin/bash
# Round 2: dK/dV kernel with double-buffered S^T at head dim 64 - parity, then A/B at config 4.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_backward_gpu.py tests/test_full_size_gpu.py tests/test_varlen_headdim_gpu.py tests/test_bias_dropout_gpu.py tests/test_triton_golden_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/dbl_tests.log 2>&1
echo "backward tests exit $?"; tail -5 gpurun_out/dbl_tests.log
for rep in 1 2; do
  echo "== before: $(FA2_SM100_LIB=variants/lib_nostage.so timeout 300 python scripts/kernel_times.py 2 16 16384 64 0 10 2>&1 | tail -1)"
  echo "== double-buffered: $(timeout 300 python scripts/kernel_times.py 2 16 16384 64 0 10 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/dbl_times.txt
echo "== causal D=64 S=4096: $(timeout 300 python scripts/kernel_times.py 4 32 4096 64 1 20 2>&1 | tail -1)" | tee -a gpurun_out/dbl_times.txt
echo "== before: $(FA2_SM100_LIB=variants/lib_nostage.so timeout 300 python scripts/kernel_times.py 4 32 4096 64 1 20 2>&1 | tail -1)" | tee -a gpurun_out/dbl_times.txt
