"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page), with the dominant stall reasons.
Usage: python scripts/ncu_hot.py report.ncu-rep [topN] [context] [launch index in the report]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launch = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
nxt = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Address"), len(rows))  # a second table follows
body = [r for r in rows[hi + 1:nxt] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
samp = col["# Samples"]
stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[samp]) for r in body)
print(f"total samples {total}, instructions {len(body)}")
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp]))[:topn]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), h[6:]) for h, c in stall_cols if int(r[c]) > 0), reverse=True)[:3]
    for j in range(max(0, i - ctx), i):
        print(f"      {j:5d}            {body[j][col['Source']].strip()[:110]}")
    print(f"{i:5d} {int(r[samp]):7d} {100.0 * int(r[samp]) / total:5.1f}%  {r[col['Source']].strip()[:100]:100s} {st}")
