#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -q -x --timeout 300 -k "above_128 or head_dim_256" > gpurun_out/d256_tests.log 2>&1
echo "d256 bwd tests exit $?"; tail -12 gpurun_out/d256_tests.log
timeout 900 python -m pytest tests/test_bias_dropout_gpu.py tests/test_varlen_headdim_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/d256_tests2.log 2>&1
echo "bias/varlen tests exit $?"; tail -8 gpurun_out/d256_tests2.log
timeout 1200 python -m pytest tests/test_reference_grid_gpu.py -m gpu -q --timeout 600 -k "207 or 256" > gpurun_out/d256_grid.log 2>&1
echo "reference grid (207, 256) exit $?"; tail -8 gpurun_out/d256_grid.log
