"""Aggregates an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line:
executed warp instructions and stall samples.  Usage: python tools/ncu_lines.py src.csv [topN]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header rows: the file is a sequence of blocks (one per source file) each with its own header
agg = defaultdict(lambda: [0, 0, ""])
cur_file, hdr = None, None
total = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {name: i for i, name in enumerate(r)}
        # two "Source" columns: first is CUDA-C text, second is SASS
        continue
    if hdr is None or len(r) < 10:
        continue
    try:
        line = r[0]
        inst = int(float(r[hdr["Instructions Executed"]] or 0))
        samp = int(float(r[hdr["# Samples"]] or 0))
    except Exception:
        continue
    key = (cur_file, line)
    agg[key][0] += inst
    agg[key][1] += samp
    if not agg[key][2]:
        agg[key][2] = r[1][:90]
    total += inst
tot_s = sum(v[1] for v in agg.values())
print(f"total warp instructions {total}, samples {tot_s}")
for (f, line), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:12d} {100.0 * v[0] / max(total, 1):5.1f}%  samp {100.0 * v[1] / max(tot_s, 1):5.1f}%  {f}:{line}  {v[2]}")
