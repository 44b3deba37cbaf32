"""Compact SASS stream of a kernel with stall samples per window.  Usage: ncu_stream.py sass.csv [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = {n: i for i, n in enumerate(rows[hi])}
def code(o):
    for k, c in (("DFMA", "F"), ("DMUL", "M"), ("DADD", "A"), ("LDS.128", "L"), ("LDS", "l"), ("STS", "S"), ("SHFL", "H"), ("MUFU", "U"),
                 ("WARPSYNC", "|"), ("NOP", "|"), ("BRA", "B"), ("BSSY", "B"), ("BSYNC", "B"), ("LDGSTS", "G"), ("LDG", "G"), ("STG", "W"), ("DSETP", "D")):
        if o.startswith(k): return c
    return "."
ops = []
for r in rows[hi + 1:]:
    if len(r) < 20: continue
    o = [t for t in r[1].split() if not t.startswith("@")][0]
    f = lambda k: int(float(r[hdr[k]] or 0)) if k in hdr else 0
    ops.append((code(o), f("# Samples"), f("Instructions Executed"), f("stall_wait"), f("stall_short_sb"), f("stall_long_sb"), f("stall_math"), f("stall_not_selected"), f("stall_no_inst")))
tot = sum(o[1] for o in ops); toti = sum(o[2] for o in ops)
print("instr", len(ops), "samples", tot)
for i in range(0, len(ops), W):
    w = ops[i:i + W]
    s = sum(o[1] for o in w); ins = sum(o[2] for o in w)
    print(f"{i:5d} {''.join(o[0] for o in w):{W}s} t {100*s/tot:4.1f}% i {100*ins/toti:4.1f}% wait {100*sum(o[3] for o in w)/tot:4.1f} ssb {100*sum(o[4] for o in w)/tot:4.1f} lsb {100*sum(o[5] for o in w)/tot:4.1f} math {100*sum(o[6] for o in w)/tot:4.1f}")
