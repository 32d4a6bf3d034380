#!/bin/bash
# One gpurun call: new tests first (all failures shown), then the whole GPU suite, smoke and a bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_triton_golden_gpu.py tests/test_bias_dropout_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/new_tests.log 2>&1
echo "new tests exit $?" | tee -a gpurun_out/new_tests.log
tail -25 gpurun_out/new_tests.log
if [ "$1" != "quick" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/all_tests.log 2>&1
  echo "all tests exit $?" | tee -a gpurun_out/all_tests.log
  tail -8 gpurun_out/all_tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
fi
