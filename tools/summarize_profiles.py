"""Turns the .ncu-rep captures in gpurun_out/ into the small text/JSON summaries committed under profiles/ and
regenerates profiles/roofline_traffic.json, which bench.py reads for `roofline.traffic`, `roofline.fp64_inst_frac`
and `roofline.max_algorithmic_frac_at_full_pipe` (keyed by the library's kernel name).

    python tools/summarize_profiles.py r02c           # captures gpurun_out/<prefix>_u_band_*.ncu-rep

Every capture: `ncu --set full --clock-control none --import-source on`, one launch of tools/kbench.py's kernel."""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
prefix = sys.argv[1] if len(sys.argv) > 1 else "r02c"
out_tag = sys.argv[2] if len(sys.argv) > 2 else "r02"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def num(d, k):
    v, u = d[k]
    v = float(v.replace(",", ""))
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


def opcodes(rep, sets):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True).stdout.decode("utf-8", "replace")
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = {n: i for i, n in enumerate(rows[hi])}
    agg = defaultdict(lambda: [0, 0, 0])
    for r in rows[hi + 1:]:
        if len(r) < 20:
            continue
        o = [t for t in r[1].split() if not t.startswith("@")][0]
        o = ".".join(o.split(".")[:2])
        a = agg[o]
        a[0] += int(float(r[hdr["Instructions Executed"]] or 0))
        a[1] += int(float(r[hdr["L1 Wavefronts Shared"]] or 0))
        a[2] += int(float(r[hdr["L1 Wavefronts Shared Ideal"]] or 0))
    fp64 = sum(a[0] for o, a in agg.items() if re.match(r"D(FMA|MUL|ADD|SETP|MNMX)", o)) / sets
    lines = [f"{o:18s} warp-inst/set {a[0] / sets:8.2f}   shared wavefronts/set {a[1] / sets:8.2f} (ideal {a[2] / sets:8.2f})"
             for o, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]]
    return fp64, lines


def summary(rep, tag, sets, what):
    d = raw(rep)
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(d[k][0]), 3)
              for k in d if re.search(r"smsp__average_warps_issue_stalled.*_per_issue_active", k) and float(d[k][0]) > 0.05}
    dur = float(d["gpu__time_duration.sum"][0]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}[d["gpu__time_duration.sum"][1]]
    fp64_inst, op_lines = opcodes(rep, sets)
    s = {
        "kernel": d["Kernel Name"][0] if "Kernel Name" in d else tag,
        "capture": f"ncu --set full --clock-control none --import-source on, one launch, {what}",
        "gpu_time_ms_under_ncu": dur,
        "dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum"),
        "dram_bytes_per_launch": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"),
        "l2_sector_hit_rate_pct": float(d["lts__t_sector_hit_rate.pct"][0]),
        "fp64_pipe_pct_of_peak_active": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "fp64_warp_inst_per_set": fp64_inst,
        "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "l1tex_data_pipe_lsu_wavefronts_pct": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0]),
        "warp_instructions_per_set": float(d["smsp__inst_executed.sum"][0]) / sets,
        "shared_wavefronts_per_set": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0]) / sets,
        "registers_per_thread": float(d["launch__registers_per_thread"][0]),
        "warps_active_per_scheduler": float(d["smsp__warps_active.avg.per_cycle_active"][0]),
        "l1_sector_hit_rate_pct": float(d["l1tex__t_sector_hit_rate.pct"][0]),
        "stall_reasons_per_issue": stalls,
    }
    json.dump(s, open(os.path.join(pr, f"{out_tag}_{tag}_summary.json"), "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(pr, f"{out_tag}_{tag}_details.txt"), "w").write("\n".join(l for l in det.splitlines() if l.strip()))
    open(os.path.join(pr, f"{out_tag}_{tag}_opcodes.txt"), "w").write(
        f"SASS opcode histogram of {s['kernel']} ({what}); fp64-pipe warp instructions per set: {fp64_inst:.1f}\n" + "\n".join(op_lines) + "\n")
    return s


CAPTURES = [  # (file tag, sets in the launch, library kernel name, description)
    ("u_band_closed_P31_D2_nu15", 1e6, "u_band<G=8,P=31,D=2,closed>", "n = 1e6 sets, m = 30, d = 2, Matern nu = 1.5"),
    ("u_band_general_P31_D2_nu08", 1e6, "u_band<G=8,P=31,D=2,general>", "n = 1e6 sets, m = 30, d = 2, Matern nu = 0.8"),
    ("u_band_closed_P41_D3_nu15", 1e6, "u_band<G=16,P=41,D=3,closed>", "n = 1e6 sets, m = 40, d = 3, Matern nu = 1.5"),
    ("u_band_closed_P31_D2_nu15_n8e6", 8e6, None, "n = 8e6 sets, m = 30, d = 2, Matern nu = 1.5 (locality layer on)"),
]
kernels = {}
for tag, sets, kname, what in CAPTURES:
    rep = os.path.join(go, f"{prefix}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    s = summary(rep, tag, sets, what)
    print(tag, {k: s[k] for k in ("gpu_time_ms_under_ncu", "dram_bytes_per_launch", "fp64_pipe_pct_of_peak_active",
                                  "fp64_warp_inst_per_set", "l1tex_data_pipe_lsu_wavefronts_pct", "shared_wavefronts_per_set")})
    if kname:
        kernels[kname] = {"source": f"profiles/{out_tag}_{tag}_summary.json", "sets_per_launch": int(sets) - 1,
                          "dram_bytes_per_launch": s["dram_bytes_per_launch"],
                          "fp64_warp_inst_per_set": s["fp64_warp_inst_per_set"],
                          "fp64_pipe_pct_of_peak_active": s["fp64_pipe_pct_of_peak_active"],
                          "gpu_time_ms_under_ncu": s["gpu_time_ms_under_ncu"]}
path = os.path.join(pr, "roofline_traffic.json")
old = {}
if os.path.exists(path):
    try:
        old = json.load(open(path)).get("kernels", {})
    except Exception:
        old = {}
old.update(kernels)
json.dump({"doc": "per-kernel evidence from the committed ncu captures (tools/summarize_profiles.py); bench.py looks its dominant "
                  "kernel up by name", "kernels": old}, open(path, "w"), indent=1)
