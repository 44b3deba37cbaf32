"""Kernel micro-benchmark for development: times the set kernel (CUDA events inside the library)
for several covariances at n rows, m neighbours; optional parity check against the oracle on a
row sample.  GPV_LIB_PATH selects the library build.  Usage: python tools/kbench.py [n] [m] [d]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import gpvecchia_b200 as G  # noqa: E402
from gpvecchia_b200 import harness as H  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 30
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
check = os.environ.get("KBENCH_CHECK", "1") == "1"
locs = H.make_locs(n, d, stream=2)
revNN = H.ordered_nn_gpu(locs, m)
revCond = np.zeros(revNN.shape, dtype=np.int32)
revCond[revNN == 0] = np.iinfo(np.int32).min
revCond[:, -1] = 1
nug = H.make_nuggets(n, stream=2)
rng_ = H.default_range(n, d)
h = G.UHandle(locs, revNN, revCond, obs=np.ones(n, dtype=np.int32))
dev = torch.device("cuda", 0)
d_nug = torch.from_numpy(nug).to(dev)
d_out = torch.empty(n * (m + 1), dtype=torch.float64, device=dev)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
covs = [("nu0.5", "matern", [1.0, rng_, 0.5]), ("nu1.5", "matern", [1.0, rng_, 1.5]),
        ("nu2.5", "matern", [1.0, rng_, 2.5]), ("gen0.8", "matern", [1.0, rng_, 0.8]),
        ("esqe", "esqe", [0.7, rng_, 0.3, rng_])]
res = {}
for tag, ct, cp in covs:
    cp = np.array(cp)
    ms = []
    for it in range(8):
        h.u_dev(ct, cp, d_nug.data_ptr(), d_out.data_ptr(), stream=s.cuda_stream)
        ms.append(h.last_kernel_ms())
    res[tag] = (min(ms[2:]), float(np.mean(ms[2:])))
    if check:
        import oracle as O
        torch.cuda.synchronize()
        k = 20000
        got = d_out.view(n, m + 1)[n - k:].cpu().numpy()
        rc = revCond[n - k:].astype(np.float64)
        rc[revCond[n - k:] < 0] = np.nan
        pr = O.RowsProblem(locs, revNN[n - k:], rc, n - k, nug, ct, cp)
        pr.run(O.max_threads())
        ref = pr.Lentries()
        err = float((np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max())
        res[tag] += (err,)
print(os.environ.get("GPV_LIB_PATH", "default"), h.last_kernel_name(), f"n={n} m={m} d={d}")
for tag, v in res.items():
    print(f"  {tag:8s} min {v[0]:.3f} ms  avg {v[1]:.3f} ms  {n / v[0] / 1e3:.1f} Msets/s" + (f"  err {v[2]:.1e}" if len(v) > 2 else ""))
