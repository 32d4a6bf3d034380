#!/bin/bash
# ncu evidence for the current binary: launch list of bench.py, then one full capture per hot kernel (3 launches each).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python scripts/profile_target.py > gpurun_out/target_plain.log 2>&1 || exit 1
for k in fa2_fwd_kernel fa2_bwd_dkdv_kernel fa2_bwd_dq_spill_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o gpurun_out/$k -f \
      python scripts/profile_target.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/$k.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
done
# the bias variants (forward kExt = 5, dK/dV kExtra = 1), one capture each
for k in fa2_fwd_kernel fa2_bwd_dkdv_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/bias_$k -f \
      python scripts/profile_target.py 4 32 4096 128 1 3 bias > gpurun_out/ncu_bias_$k.log 2>&1
  ncu -i gpurun_out/bias_$k.ncu-rep --page raw --csv > gpurun_out/bias_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -15
