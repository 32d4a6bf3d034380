#!/bin/bash
# generic A/B inside one call: bash scripts/gpu_ab_variants.sh "<variant names under variants/lib_<name>.so; main = the regular build>" [workload] [reps]
mkdir -p gpurun_out
W=${2:-cfg3}; R=${3:-2}
for rep in $(seq 1 $R); do
for v in $1; do
  lib=variants/lib_$v.so; if [ $v = main ]; then lib=fa2_triton_b200/libfa2_sm100.so; fi
  FA2_SM100_LIB=$lib timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-triton > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - $v <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, {k: round(d[k], 4) for k in ("value", "ms_per_step")}, "sustained", round(d["sustained"]["ms_per_step"], 4), d["sustained"]["clocks"]["sm_mhz"],
          {k[:26]: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(n, "parse failed", e, open(f"gpurun_out/ab_{n}.err").read()[-600:])
PY
done
done | tee gpurun_out/ab_last.txt
