"""Shared-memory wavefronts, executed warp instructions and stall samples per CUDA source line of one capture.
Usage: python tools/ncu_wavefronts.py rep.ncu-rep nsets [topN]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, nsets = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(txt)))
agg = defaultdict(lambda: [0, 0, 0, 0, ""])
cur, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < 20:
        continue
    try:
        inst = int(float(r[hdr["Instructions Executed"]] or 0)); samp = int(float(r[hdr["# Samples"]] or 0))
        wf = int(float(r[hdr["L1 Wavefronts Shared"]] or 0)); ideal = int(float(r[hdr["L1 Wavefronts Shared Ideal"]] or 0))
    except Exception:
        continue
    a = agg[(cur, r[0])]
    a[0] += inst; a[1] += samp; a[2] += wf; a[3] += ideal
    if not a[4]: a[4] = r[1].strip()[:100]
ti, ts, tw = (sum(v[i] for v in agg.values()) for i in (0, 1, 2))
print(f"total: {ti / nsets:.1f} warp-instr/set, {tw / nsets:.1f} shared wavefronts/set, {ts} samples")
print("by shared wavefronts:")
for (f, line), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    if v[2] == 0: break
    print(f"  wf/set {v[2] / nsets:7.2f} (ideal {v[3] / nsets:7.2f}) {100.0 * v[2] / tw:5.1f}%  inst/set {v[0] / nsets:7.2f}  samp {100.0 * v[1] / ts:5.1f}%  {f}:{line}  {v[4]}")
print("by stall samples:")
for (f, line), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  samp {100.0 * v[1] / ts:5.1f}%  inst/set {v[0] / nsets:7.2f}  {f}:{line}  {v[4]}")
