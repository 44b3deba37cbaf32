"""Key metrics + SASS opcode/wavefront histogram of one .ncu-rep.  Usage: ncu_brief.py rep nsets"""
import csv, subprocess, sys, re
from collections import defaultdict
rep, nsets = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = {h: v for h, v in zip(rows[0], rows[2])}
def g(k): return float(d[k].replace(",", ""))
print("kernel", d.get("Kernel Name"), "time", d["gpu__time_duration.sum"], rows[1][rows[0].index("gpu__time_duration.sum")])
for k in ["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
          "smsp__warps_active.avg.per_cycle_active"]:
    print(f"  {k}: {d[k]}")
print("  inst/set", g("smsp__inst_executed.sum") / nsets, " shared wf/set", g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / nsets,
      " cycles/set/SM", g("sm__cycles_elapsed.max") * 148 / nsets if "sm__cycles_elapsed.max" in d else "")
st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(d[k]), 2)
      for k in d if re.search(r"smsp__average_warps_issue_stalled.*_per_issue_active", k) and float(d[k]) > 0.05}
print("  stalls/issue", st)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = {n: i for i, n in enumerate(rows[hi])}
agg = defaultdict(lambda: [0, 0, 0])
for r in rows[hi + 1:]:
    if len(r) < 20: continue
    o = [t for t in r[1].split() if not t.startswith("@")][0]
    o = ".".join(o.split(".")[:2])
    a = agg[o]
    a[0] += int(float(r[hdr["Instructions Executed"]] or 0)); a[1] += int(float(r[hdr["L1 Wavefronts Shared"]] or 0)); a[2] += int(float(r[hdr["L1 Wavefronts Shared Ideal"]] or 0))
for o, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {o:18s} inst/set {a[0]/nsets:7.1f}  wf/set {a[1]/nsets:7.1f} ideal {a[2]/nsets:7.1f}")
