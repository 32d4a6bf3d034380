#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/flake_suite_$i.log 2>&1
  echo "suite run $i exit $? $(tail -1 gpurun_out/flake_suite_$i.log)"
  grep -h "^E \|^FAILED" gpurun_out/flake_suite_$i.log | head -8
done
