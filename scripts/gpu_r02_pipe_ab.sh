#!/bin/bash
# A/B of pipelined-backward variants, one call
mkdir -p gpurun_out
for g in 1 3 4; do FA2_SM100_PIPE_G=$g timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x --timeout 120 -k pipelined 2>&1 | tail -1; done
for rep in 1 2; do
for v in seq main_g1 main_g4 main_g8 pipeR_g4 pipeR_g8; do
  lib=fa2_triton_b200/libfa2_sm100.so; pipe=1; g=4
  case $v in seq) pipe=0;; main_g1) g=1;; main_g8) g=8;; pipeR_g4) lib=variants/lib_pipeR.so;; pipeR_g8) lib=variants/lib_pipeR.so; g=8;; esac
  FA2_SM100_LIB=$lib FA2_SM100_PIPE=$pipe FA2_SM100_PIPE_G=$g timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
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
done | tee gpurun_out/pipe_ab8.txt
