#!/bin/bash
# pipelined backward: bit-identity test, A/B against the two kernels in sequence (and the relaxed-signal variant), then one
# ncu --set full capture of the pipelined kernel (DRAM traffic is the point: do the dS tiles stay in L2?)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x --timeout 120 -k pipelined 2>&1 | tail -2
for rep in 1 2; do
for v in seq main pipeR nodiscard; do
  lib=fa2_triton_b200/libfa2_sm100.so; pipe=1
  if [ $v = seq ]; then pipe=0; fi
  if [ $v = nodiscard ]; then pipe=2; fi
  if [ $v = pipeR ]; then lib=variants/lib_pipeR.so; fi
  FA2_SM100_LIB=$lib FA2_SM100_PIPE=$pipe timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - $v <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, {k: round(d[k], 4) for k in ("value", "ms_per_step")}, "sustained", round(d["sustained"]["ms_per_step"], 4), d["sustained"]["clocks"]["sm_mhz"],
          {k[:34]: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(n, "parse failed", e, open(f"gpurun_out/bench_{n}.err").read()[-600:])
PY
done
done | tee gpurun_out/pipe_ab7.txt
for v in main nodiscard; do
  pipe=1; if [ $v = nodiscard ]; then pipe=2; fi
  FA2_SM100_PIPE=$pipe timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa2_bwd_pipe_kernel -s 1 -c 1 -o gpurun_out/r02d_pipe_$v -f \
      python scripts/profile_target.py > gpurun_out/ncu_pipe_$v.log 2>&1
  ncu -i gpurun_out/r02d_pipe_$v.ncu-rep --page raw --csv > gpurun_out/r02d_pipe_${v}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r02d_pipe_${v}_raw.csv gpurun_out/r02d_pipe_${v}_ncu_raw.txt "fa2_bwd_pipe_kernel ($v), config 3 (B=4,H=32,S=4096,D=128 causal bf16), 1 launch"
  grep "dram__bytes_read.sum \|dram__bytes_write.sum \|gpu__time_duration.sum\|sm__pipe_tensor_cycles_active.avg.pct\|lts__t_sector_hit_rate" gpurun_out/r02d_pipe_${v}_ncu_raw.txt
done
