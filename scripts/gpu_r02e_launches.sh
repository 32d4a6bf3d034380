#!/bin/bash
# ncu launch list of the final round-2 binary (kernel shares of a step), after the same command has run without ncu
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-triton --sustain-s 0 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { tail -5 gpurun_out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-triton --sustain-s 0 > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02e_launches.csv")) if len(r) > 14 and r[0].isdigit()]
t = collections.defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].split("<")[0].replace("void ", "")
    if "fa2" in name: t[name].append(float(r[14]))
tot = sum(sum(v) for v in t.values())
for k, v in t.items(): print(k, len(v), "launches, mean", round(sum(v) / len(v) / 1e6, 4), "ms, share", round(100 * sum(v) / tot, 1), "%")
PY
