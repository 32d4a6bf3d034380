"""Splits a kernel's SASS (ncu source page) at synchronisation instructions and prints, per region, the number of
instructions, warp-stall samples and the dominant stall reasons. Usage: python scripts/ncu_regions.py report.ncu-rep [min_samples]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [(h[6:], i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
KEYS = ("SYNCS", "UTCBAR", "BAR.", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCHMMA", "EXIT", "UBLKRED", "FENCE", "MEMBAR")
total = sum(int(r[col["# Samples"]]) for r in body)
print(f"total samples {total}; {len(body)} instructions")
start = 0
def flush(a, b, label):
    n = sum(int(body[i][col["# Samples"]]) for i in range(a, b))
    if n < min_s: return
    agg = collections.Counter()
    ex = 0
    for i in range(a, b):
        ex += int(body[i][col["Instructions Executed"]])
        for h, c in stall_cols: agg[h] += int(body[i][c])
    top = ", ".join(f"{k}={v}" for k, v in agg.most_common(4) if v)
    print(f"[{a:5d},{b:5d}) {b - a:4d} instr {ex:11d} exec {n:7d} samples {100.0 * n / total:5.1f}%  ends at: {label[:70]:70s} | {top}")
for i, r in enumerate(body):
    s = r[col["Source"]].strip()
    if any(k in s for k in KEYS):
        flush(start, i + 1, s)
        start = i + 1
flush(start, len(body), "end")
