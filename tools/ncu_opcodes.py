"""Dynamic opcode histogram from an `ncu --page source --csv --print-source sass` export."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = defaultdict(int); samp = defaultdict(int)
tot = 0
for r in rows:
    if r and r[0] in ("Address", "Line No"):
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or "Address" not in hdr or len(r) < 8:
        continue
    try:
        sass = r[hdr["Source"]] if "Line No" not in hdr else r[3]
        inst = int(float(r[hdr["Instructions Executed"]] or 0))
        s = int(float(r[hdr["# Samples"]] or 0))
    except Exception:
        continue
    toks = sass.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0].rstrip(";")
    agg[op] += inst; samp[op] += s; tot += inst
ts = sum(samp.values())
print("total", tot)
for op, v in sorted(agg.items(), key=lambda kv: -kv[1])[:30]:
    print(f"{op:14s} {v:12d} {100*v/tot:5.1f}%   stall samples {100*samp[op]/max(ts,1):5.1f}%")
