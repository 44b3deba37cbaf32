"""Attributes executed warp instructions and stall samples (time) of the set kernel to its phases,
from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`.  Usage: ncu_phases.py src.csv nwarps"""
import csv, re, sys, os
from collections import defaultdict
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rows = list(csv.reader(open(sys.argv[1])))
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else 5e5
lines = open(os.path.join(root, "gpvecchia_b200/csrc/u_kernels.cuh")).read().split("\n")
marks = []
for i, t in enumerate(lines, 1):
    m = re.search(r"// ---- (\d)\. (\w+)", t)
    if m: marks.append((i, "step" + m.group(1) + "_" + m.group(2)))
    for pat, name in (("deterministic block reduction", "reduce"), ("double rsqrt_pos", "rsqrt_pos"),
                      ("double sqrt_nonneg", "sqrt"), ("double exp_neg", "exp"), ("template <int KIND>\n", "cov"),
                      ("struct SetLayout", "layout"), ("double pair_r2", "pair_r2"), ("void pair_eval_store", "pair_store"),
                      ("void pair_stage", "pair_loop"), ("u_sets_kernel(const UParams q) {", "prologue"),
                      ("double cov_eval", "cov_eval"), ("double rsqrt_seed", "rsqrt_seed")):
        if pat in t: marks.append((i, name))
marks.sort()
def ph_of(l):
    name = "top"
    for i, n in marks:
        if l >= i: name = n
    return name
def num(x):
    try: return int(float(x))
    except Exception: return 0
ph = defaultdict(lambda: [0, 0]); tot = ts = 0; hdr = None; cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < 10 or not r[0].strip().isdigit(): continue
    inst = num(r[hdr["Instructions Executed"]]); s = num(r[hdr["# Samples"]])
    key = ph_of(int(r[0])) if cur == "u_kernels.cuh" else cur
    ph[key][0] += inst; ph[key][1] += s; tot += inst; ts += s
print("total inst", tot, "per warp", tot / nwarps, "samples", ts)
for k, v in sorted(ph.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:22s} inst {v[0]/nwarps:8.1f}/warp {100*v[0]/tot:5.1f}%   time {100*v[1]/ts:5.1f}%   rel.cost/inst {v[1]/max(v[0],1)*tot/ts:5.2f}")
