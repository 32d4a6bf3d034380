#!/bin/bash
# evidence of the FINAL round-2 binary: widened rows, ncu --set full of the three hot kernels at config 3
mkdir -p gpurun_out
timeout 900 python scripts/bench_widened.py > gpurun_out/r02e_widened_rows.jsonl 2> gpurun_out/widened.err; echo "widened exit $?"; tail -3 gpurun_out/widened.err
python scripts/profile_target.py > gpurun_out/target_plain.log 2>&1 || exit 1
for k in fa2_fwd2_kernel fa2_bwd_dkdv_kernel fa2_bwd_dq_spill_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o gpurun_out/r02e_$k -f \
      python scripts/profile_target.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/r02e_$k.ncu-rep --page raw --csv > gpurun_out/r02e_${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r02e_${k}_raw.csv gpurun_out/r02e_${k}_ncu_raw.txt "$k, final round-2 binary (tail-balanced order), config 3 (B=4,H=32,S=4096,D=128 causal bf16), 2 launches"
  grep "gpu__time_duration.sum \|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active\|dram__bytes_read.sum \[\|dram__bytes_write.sum \[" gpurun_out/r02e_${k}_ncu_raw.txt
done
wc -l gpurun_out/r02e_widened_rows.jsonl
