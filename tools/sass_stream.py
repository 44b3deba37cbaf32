"""Compact opcode stream of one kernel from the built library.  Usage: sass_stream.py <mangled-substring> [lib]"""
import subprocess, sys, re
lib = sys.argv[2] if len(sys.argv) > 2 else "gpvecchia_b200/libgpvecchia_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ops = None, []
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); continue
    if cur and sys.argv[1] in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m: ops.append(m.group(1).split())
def code(t):
    o = t[1] if t[0].startswith("@") else t[0]
    for k, c in (("DFMA", "F"), ("DMUL", "M"), ("DADD", "A"), ("LDS.128", "L"), ("LDS", "l"), ("STS", "S"), ("SHFL", "H"), ("MUFU", "U"),
                 ("WARPSYNC", "|"), ("NOP", "|"), ("BRA", "B"), ("BSSY", "B"), ("BSYNC", "B"), ("LDGSTS", "G"), ("LDG", "G"), ("STG", "W"),
                 ("DSETP", "D"), ("CALL", "C"), ("RET", "R"), ("LDL", "x"), ("STL", "X")):
        if o.startswith(k): return c
    return "."
s = "".join(code(t) for t in ops)
print(len(s), "instructions")
for i in range(0, len(s), 150): print(f"{i:5d} {s[i:i+150]}")
