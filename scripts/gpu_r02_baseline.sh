#!/bin/bash
# Round 2, first call: the reference's Triton kernels beside ours (timing + bf16 parity), then the reference's own
# pytest files through our entry point.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/ls_reference.txt 2>&1; ls baseline/_ref >> gpurun_out/ls_reference.txt 2>&1
timeout 1500 python scripts/triton_baseline.py --out gpurun_out/r02_triton_baseline.json > gpurun_out/triton_baseline.log 2>&1
echo "triton_baseline exit $?"; tail -c 3000 gpurun_out/triton_baseline.log
timeout 1500 python scripts/run_reference_tests.py --tb=line -q > gpurun_out/ref_tests_ours.log 2>&1
echo "reference tests (ours) exit $?"; tail -5 gpurun_out/ref_tests_ours.log
timeout 600 python scripts/run_reference_tests.py --impl triton --tb=line -q -k "128 and 1023 and dtype1" > gpurun_out/ref_tests_triton_control.log 2>&1
echo "reference tests (triton control subset) exit $?"; tail -5 gpurun_out/ref_tests_triton_control.log
