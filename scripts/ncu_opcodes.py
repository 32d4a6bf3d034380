"""Executed warp instructions by opcode from an .ncu-rep (source page). Usage: python scripts/ncu_opcodes.py report.ncu-rep [topN]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter()
for r in body:
    op = r[col["Source"]].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o if o.startswith("MUFU") else o.split(".")[0]
    agg[o] += int(r[col["Instructions Executed"]])
tot = sum(agg.values())
print(f"total warp instructions {tot}")
for o, c in agg.most_common(topn): print(f"{o:14s} {c:12d} {100 * c / tot:5.1f}%")
