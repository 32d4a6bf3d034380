#!/bin/bash
# Round 2: whole GPU suite, smoke, bench line, power probe.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/all_tests.log 2>&1
echo "all tests exit $?"; tail -6 gpurun_out/all_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "frac_of_measured_burst_peak")})
    print("sustained", {k: d["sustained"][k] for k in ("value", "ms_per_step", "steps")}, d["sustained"]["clocks"])
    print("kernels", {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()}, "sum", d["roofline"]["sum_of_kernels_ms"])
    print("clocks", d["clocks"]); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("triton", d["triton_baseline"]); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python scripts/power_probe.py > gpurun_out/power_probe.txt 2>&1; cat gpurun_out/power_probe.txt
