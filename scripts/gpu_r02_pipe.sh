#!/bin/bash
# Round 2: pipelined backward (one persistent kernel, CTAs switch between dK/dV items and ready dQ items): bit-identity
# test first, then the backward tests, then A/B inside one call (FA2_SM100_PIPE=0 two kernels in sequence, 1 pipelined with
# L2 discard, 2 pipelined without discard).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x --timeout 120 -k pipelined > gpurun_out/pipe_tests.log 2>&1
echo "pipelined test exit $?"; tail -15 gpurun_out/pipe_tests.log
if ! grep -q "passed" gpurun_out/pipe_tests.log || grep -q "failed" gpurun_out/pipe_tests.log; then exit 1; fi
for rep in 1 2; do
for n in 0 1 2; do
  FA2_SM100_PIPE=$n timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_pipe$n.json 2> gpurun_out/bench_pipe$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_pipe{n}.json").read().strip().splitlines()[-1])
    print("pipe", n, {k: round(d[k], 4) for k in ("value", "ms_per_step")}, "sustained", round(d["sustained"]["ms_per_step"], 4), d["sustained"]["clocks"]["sm_mhz"],
          {k[:34]: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(n, "parse failed", e, open(f"gpurun_out/bench_pipe{n}.err").read()[-600:])
PY
done
done | tee gpurun_out/pipe_ab.txt
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pipe_tests2.log 2>&1
echo "backward + full-size tests exit $?"; tail -5 gpurun_out/pipe_tests2.log
