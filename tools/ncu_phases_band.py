"""Phase attribution of the band kernel from an ncu report with imported source:
    ncu -i X.ncu-rep --page source --csv --print-source cuda      > src.csv     (source text as captured)
    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > sass.csv    (per-line counters)
    python tools/ncu_phases_band.py sass.csv src.csv
Sums stall samples (time), executed warp instructions and shared-memory wavefronts per phase; phase boundaries
are the `// ---- N.` markers of u_band.cuh as embedded in the report, so it works on captures of older builds."""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
cur=None; hdr=None
agg=defaultdict(lambda:[0,0,0,0,0,0])  # samples, inst, wavefronts, short_sb, wait, math
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; hdr=None; continue
    if r and r[0]=="Line No": hdr={n:i for i,n in enumerate(r)}; continue
    if hdr is None or len(r)<10 or not r[0].strip().isdigit(): continue
    def g(n):
        try: return float(r[hdr[n]])
        except Exception: return 0.0
    k=(cur,int(r[0]))
    a=agg[k]; a[0]+=g("# Samples"); a[1]+=g("Instructions Executed"); a[2]+=g("L1 Wavefronts Shared"); a[3]+=g("stall_short_sb"); a[4]+=g("stall_wait"); a[5]+=g("stall_math")
# source text of u_band.cuh at capture time
src={}
for r in csv.reader(open(sys.argv[2])):
    if len(r)>=2 and r[0].isdigit(): src[int(r[0])]=",".join(r[1:])
# find phase boundaries from comment markers in captured source
bounds=[]
for ln in sorted(src):
    tx=src[ln]
    m=re.search(r"// ---- (\d)\.(/\d\.)? ",tx)
    if m: bounds.append((ln,"step"+m.group(1)))
    if "pair stage: point i evaluates" in tx: bounds.append((ln,"pair_loop"))
    if "u_band_kernel(const UParams q)" in tx: bounds.append((ln,"prologue"))
    if "input pipeline (cp.async" in tx: bounds.append((ln,"fetch/gather lambdas"))
    if "int bsel = 0;" in tx: bounds.append((ln,"pipeline start"))
    if "deterministic block reduction" in tx: bounds.append((ln,"reduce"))
    if "build_store_table_band" in tx and "void" in tx: bounds.append((ln,"store table"))
bounds.sort()
def ph(f,l):
    if f=="u_kernels.cuh": return "cov math (u_kernels.cuh)"
    if f!="u_band.cuh": return "hdr:"+f
    name="top"
    for b,n in bounds:
        if l>=b: name=n
    return name
P=defaultdict(lambda:[0,0,0,0,0,0])
for (f,l),a in agg.items():
    p=P[ph(f,l)]
    for i in range(6): p[i]+=a[i]
ts=sum(p[0] for p in P.values()); ti=sum(p[1] for p in P.values()); tw=sum(p[2] for p in P.values())
print("bounds",bounds)
print(f"{'phase':32s} time%  inst%  wavefr%  short_sb% wait% math%")
for k,p in sorted(P.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:32s} {100*p[0]/ts:5.1f}  {100*p[1]/ti:5.1f}  {100*p[2]/max(tw,1):5.1f}   {100*p[3]/max(p[0],1):5.1f}   {100*p[4]/max(p[0],1):5.1f}  {100*p[5]/max(p[0],1):5.1f}")
