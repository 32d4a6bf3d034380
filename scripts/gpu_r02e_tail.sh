timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -1
bash scripts/gpu_ab_variants.sh "notail main" cfg3 2
bash scripts/gpu_ab_variants.sh "notail main" cfg2 2
bash scripts/gpu_tail_ncu.sh
