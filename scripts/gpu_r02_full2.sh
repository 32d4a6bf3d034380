#!/bin/bash
# Round 2 (second half): whole GPU suite, smoke, bench lines for configs 3 / 2 / 4 the way the driver runs them.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench.err
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/all_tests.log 2>&1
echo "all tests exit $?"; tail -4 gpurun_out/all_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
done
python - <<'PY'
import json
for f in ("bench", "bench_cfg2", "bench_cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: round(d[k], 4) for k in ("value", "ms_per_step", "frac_of_measured_burst_peak")})
        print("  sustained", {k: round(d["sustained"][k], 4) for k in ("value", "ms_per_step")}, d["sustained"]["clocks"]["sm_mhz"], d["sustained"]["clocks"]["reasons"])
        print("  kernels", {k[:26]: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()}, "sum", round(d["roofline"]["sum_of_kernels_ms"], 4))
        print("  clocks", d["clocks"])
        if d.get("triton_baseline"): print("  triton", d["triton_baseline"].get("value"), d["triton_baseline"].get("speedup_ours_over_triton"))
        if "e2e" in d and d["e2e"]: print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "parse failed", e)
PY
