#!/bin/bash
# Round 2 ncu evidence for the current binary: launch list of bench.py, then one `--set full` capture per hot kernel
# (config 3: forward, dK/dV + spill, dQ from spill; config 4 / head dim 64: forward, dK/dV, recomputing dQ).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-triton --sustain-s 0 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { tail -5 gpurun_out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-triton --sustain-s 0 > gpurun_out/ncu_launches.log 2>&1
python scripts/profile_target.py > gpurun_out/target_plain.log 2>&1 || exit 1
for k in fa2_fwd2_kernel fa2_bwd_dkdv_kernel fa2_bwd_dq_spill_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o gpurun_out/r02_$k -f \
      python scripts/profile_target.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r02_${k}_raw.csv gpurun_out/r02_${k}_ncu_raw.txt "$k, config 3 (B=4,H=32,S=4096,D=128 causal bf16), 2 launches"
done
python scripts/profile_target.py 2 16 16384 64 0 2 > gpurun_out/target_cfg4.log 2>&1 || exit 1
for k in fa2_fwd2_kernel fa2_bwd_dkdv_kernel fa2_bwd_dq_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_cfg4_$k -f \
      python scripts/profile_target.py 2 16 16384 64 0 2 > gpurun_out/ncu_cfg4_$k.log 2>&1
  ncu -i gpurun_out/r02_cfg4_$k.ncu-rep --page raw --csv > gpurun_out/r02_cfg4_${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r02_cfg4_${k}_raw.csv gpurun_out/r02_cfg4_${k}_ncu_raw.txt "$k, config 4 (B=2,H=16,S=16384,D=64 non-causal bf16), 1 launch"
done
ls -la gpurun_out | grep r02_ | tail -20
