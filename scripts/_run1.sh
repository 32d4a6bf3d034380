mkdir -p gpurun_out
python scripts/kernel_profile.py True 4 32 4096 128 1 50 > gpurun_out/kprof50.txt 2>&1; cat gpurun_out/kprof50.txt | tail -12
python scripts/kernel_profile.py True 4 32 4096 128 1 10 > gpurun_out/kprof10.txt 2>&1; cat gpurun_out/kprof10.txt | tail -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_triton_golden_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -8 gpurun_out/memcheck.log
