"""Turns the .ncu-rep captures in gpurun_out/ into the small text/JSON summaries committed under profiles/."""
import csv, json, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}

def num(d, k):
    v, u = d[k]
    v = float(v.replace(",", ""))
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return v * mult

def summary(rep, tag, sets):
    d = raw(rep)
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(d[k][0]), 3)
              for k in d if re.search(r"smsp__average_warps_issue_stalled.*_per_issue_active", k) and float(d[k][0]) > 0.05}
    dur = float(d["gpu__time_duration.sum"][0]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}[d["gpu__time_duration.sum"][1]]
    s = {
        "kernel": d["Kernel Name"][0] if "Kernel Name" in d else tag,
        "capture": "ncu --set full --clock-control none --import-source on, one launch, n = 1e6 sets, m = 30, d = 2",
        "gpu_time_ms_under_ncu": dur,
        "dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum"),
        "dram_bytes_per_launch": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"),
        "fp64_pipe_pct_of_peak_active": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "l1tex_data_pipe_lsu_wavefronts_pct": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0]),
        "warp_instructions_per_set": float(d["smsp__inst_executed.sum"][0]) / sets,
        "shared_wavefronts_per_set": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0]) / sets,
        "shared_bank_conflicts_ld": float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"][0]),
        "shared_bank_conflicts_st": float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"][0]),
        "registers_per_thread": float(d["launch__registers_per_thread"][0]),
        "achieved_occupancy_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
        "warps_active_per_scheduler": float(d["smsp__warps_active.avg.per_cycle_active"][0]),
        "l1_sector_hit_rate_pct": float(d["l1tex__t_sector_hit_rate.pct"][0]),
        "stall_reasons_per_issue": stalls,
    }
    json.dump(s, open(os.path.join(pr, f"r01_{tag}_summary.json"), "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(pr, f"r01_{tag}_details.txt"), "w").write("\n".join(l for l in det.splitlines() if l.strip()))
    return s

c = summary(os.path.join(go, "prof_closed.ncu-rep"), "u_band_closed_P31_D2_nu15", 1e6)
g = summary(os.path.join(go, "prof_general.ncu-rep"), "u_band_general_P31_D2_nu08", 1e6)
json.dump({"kernel": c["kernel"], "source": "profiles/r01_u_band_closed_P31_D2_nu15_summary.json",
           "dram_bytes_per_launch": c["dram_bytes_per_launch"],
           "fp64_pipe_pct_of_peak_active": c["fp64_pipe_pct_of_peak_active"]},
          open(os.path.join(pr, "roofline_traffic.json"), "w"), indent=1)
# launch list: per-kernel totals and shares of the bench command
rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki])[:70]
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
allms = sum(v[1] for v in tot.values())
with open(os.path.join(pr, "r01_launch_list_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras\n")
    f.write("(input generation + handle creation + 6 device-resident steps + 12 end-to-end steps; cold-cache, serialised)\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[0]:5d} launches {v[1]:10.3f} ms {100 * v[1] / allms:6.2f}%  {k}\n")
import shutil
shutil.copy(os.path.join(go, "launches.csv"), os.path.join(pr, "r01_launches_bench_steps3.csv"))
print(open(os.path.join(pr, "r01_launch_list_summary.txt")).read())
print(json.dumps(c, indent=1)[:1500])
print({k: g[k] for k in ("gpu_time_ms_under_ncu", "fp64_pipe_pct_of_peak_active", "l1tex_data_pipe_lsu_wavefronts_pct", "warp_instructions_per_set", "l1_sector_hit_rate_pct")})
