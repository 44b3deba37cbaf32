"""Full-shard parity check of bench.py's per-rank path on ONE GPU: for every rank of a `world`-rank run of a workload,
build the rank's inputs, run the device-resident call and compare EVERY row of the shard with the oracle.
Usage: python tools/debug_parity.py [workload] [world] [n_total]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench as B
import oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = dict(B.WORKLOADS[name])
n_total = int(sys.argv[3]) if len(sys.argv) > 3 else (wl["n"] * world if wl["scaling"] == "weak" else wl["n"])
worst = 0.0
for rank in range(world):
    R = B.Runner(name, wl, n_total, world, rank, 0)
    R.step_dev()
    torch.cuda.synchronize()
    got = R.d_out.view(R.nrows, R.p).cpu().numpy()
    pb = R.pb
    t0 = time.time()
    pr = O.RowsProblem(pb["locs"], pb["revNN"], pb["revCond"], 0, pb["nug_all"], wl["covType"], pb["covparms"])
    nf = pr.run(B.host_threads(), mode=1)
    ref = pr.Lentries()
    scale = np.abs(ref).max(axis=1, keepdims=True); scale[scale == 0] = 1.0
    err = np.abs(got - ref) / scale
    rowerr = err.max(axis=1)
    patt = (got == 0) != (ref == 0)
    badrows = np.nonzero(patt.any(axis=1) | ~(rowerr < 1e-10))[0]
    print(f"rank {rank}: rows [{pb['rb']},{pb['re']}) max err {np.nanmax(rowerr):.2e}, bad rows {badrows.size}, oracle nfail {nf}, {time.time() - t0:.1f}s", flush=True)
    for r in badrows[:5]:
        print("   row", int(pb["rb"] + r), "n0", int((pb["revNN"][r] != 0).sum()), "got", got[r][:4], got[r][-2:], "ref", ref[r][:4], ref[r][-2:], "ids", pb["revNN"][r][:3], pb["revNN"][r][-2:])
    worst = max(worst, float(np.nanmax(rowerr)))
    R.close()
    del R
    torch.cuda.empty_cache()
print("worst", worst)
