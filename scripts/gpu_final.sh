#!/bin/bash
# Final round check on one GPU: whole GPU suite, smoke, the default bench line, the reference arm, the other configs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/final_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err; echo "ref exit $?"
for w in cfg2 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/final_bench_$w.json 2>> gpurun_out/final_bench.err; echo "$w exit $?"; done
python - <<'PY'
import json
for n in ("final_bench", "final_bench_reference", "final_bench_cfg2", "final_bench_cfg4", "final_bench_cfg5"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 3), "ms", (d.get("e2e") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(n, "ERR", e)
PY
