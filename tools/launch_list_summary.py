"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_list_summary.py gpurun_out/r02_launches.csv "<command that was profiled>" > profiles/r02_launch_list_summary.txt"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    t = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 300: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("(input generation + handle creation + 6 device-resident steps + e2e and likelihood steps; cold-cache, serialised: shares, not absolutes)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:5d} launches {t:10.3f} ms {100 * t / tot:6.2f}%  {k[:90]}")
