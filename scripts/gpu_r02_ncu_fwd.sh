#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_fwd.py > gpurun_out/profile_fwd_plain.log 2>&1 || { tail -5 gpurun_out/profile_fwd_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:fa2_fwd2_kernel -s 1 -c 3 -o gpurun_out/r02_fwd2 -f \
    python scripts/profile_fwd.py > gpurun_out/ncu_fwd2.log 2>&1
tail -3 gpurun_out/ncu_fwd2.log
ls -la gpurun_out/*.ncu-rep
