#!/bin/bash
# tail-balanced order vs plain head-major order under ncu: duration and DRAM traffic per kernel (same box, same call)
mkdir -p gpurun_out
for v in notail main; do
  lib=variants/lib_$v.so; if [ $v = main ]; then lib=fa2_triton_b200/libfa2_sm100.so; fi
  FA2_SM100_LIB=$lib ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fa2_ -c 12 --csv --log-file gpurun_out/tail_ncu_$v.csv \
      python scripts/profile_target.py > gpurun_out/tail_ncu_$v.log 2>&1
  python - $v <<'PY'
import csv, sys, collections
v = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/tail_ncu_{v}.csv")) if len(r) > 14 and r[0].isdigit()]
t = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    name = r[4].split("(")[0].split("<")[0].replace("void fa2::", "")
    t[name][r[12]].append(float(r[14]))
for k, m in t.items():
    print(v, k, {mm: round(sum(x) / len(x), 4) for mm, x in m.items()})
PY
done
