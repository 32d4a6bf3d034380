#!/bin/bash
# A/B of the tail-balanced CTA order (last heads dealt block-major): variants/lib_notail.so (off), lib_tail16.so (16 heads),
# the regular build (32 heads); tests first
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2
for rep in 1 2; do
for v in notail tail16 main; do
  lib=variants/lib_$v.so; if [ $v = main ]; then lib=fa2_triton_b200/libfa2_sm100.so; fi
  for w in cfg3 cfg2; do
  FA2_SM100_LIB=$lib timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-triton > gpurun_out/bench_${v}_$w.json 2> gpurun_out/bench_${v}_$w.err
  python - ${v}_$w <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, {k: round(d[k], 4) for k in ("value", "ms_per_step")}, "sustained", round(d["sustained"]["ms_per_step"], 4), d["sustained"]["clocks"]["sm_mhz"],
          {k[:26]: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(n, "parse failed", e, open(f"gpurun_out/bench_{n}.err").read()[-600:])
PY
  done
done
done | tee gpurun_out/tail_ab.txt
