"""What an R session gets: gpv_u_values_packed into PAGEABLE memory (every R vector is), reused or freshly allocated per
call (Rf_allocVector gives untouched pages: the first write faults them in), next to the page-locked buffer of bench.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import gpvecchia_b200 as G
from gpvecchia_b200 import harness as H

n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 30
locs = H.make_locs(n, 2, stream=2)
NN = H.rev(H.ordered_nn_gpu(locs, m)).astype(np.int32)
Cond = np.zeros_like(NN, dtype=np.int32)
nug = H.make_nuggets(n, stream=2)
cp = [1.0, H.default_range(n, 2), 1.5]
with G.UHandle(locs, NN, Cond, obs=np.ones(n, dtype=bool)) as h:
    total = h.packed_len + 2 * n
    pinned = torch.empty(total, dtype=torch.float64).pin_memory().numpy()
    h.values_packed("matern", cp, nug, nug, out=pinned)
    ref = pinned.copy()
    reused = np.empty(total)
    K = 12
    for name, mk in (("pinned", lambda: pinned), ("pageable, reused", lambda: reused), ("pageable, fresh per call", lambda: np.empty(total))):
        for _ in range(2):
            o = mk(); h.values_packed("matern", cp, nug, nug, out=o)
        t0 = time.perf_counter()
        for _ in range(K):
            o = mk(); h.values_packed("matern", cp, nug, nug, out=o)
        dt = (time.perf_counter() - t0) / K
        assert np.array_equal(o, ref)
        print(f"{name:26s} {dt * 1e3:8.2f} ms  {n / dt / 1e6:7.1f} Msets/s  {total * 8 / dt / 1e9:6.1f} GB/s", flush=True)
    # the call the R createU() of r_shim/ makes: dgCMatrix@x
    _, nnz, _ = h.csc_dims()
    xref, _, _ = h.values_csc("matern", cp, nug, nug)
    for name, mk in (("csc pageable, reused", lambda b=np.empty(nnz): b), ("csc pageable, fresh", lambda: np.empty(nnz))):
        for _ in range(2):
            o = mk(); h.values_csc("matern", cp, nug, nug, out=o)
        t0 = time.perf_counter()
        for _ in range(K):
            o = mk(); h.values_csc("matern", cp, nug, nug, out=o)
        dt = (time.perf_counter() - t0) / K
        assert np.array_equal(o, xref)
        print(f"{name:26s} {dt * 1e3:8.2f} ms  {n / dt / 1e6:7.1f} Msets/s  {nnz * 8 / dt / 1e9:6.1f} GB/s", flush=True)
    # the stateless reference-name route: Lentries (N x p, column-major) + Zentries
    t0 = time.perf_counter()
    for _ in range(3):
        r = h.U_NZentries("matern", cp, nug, nug)
    dt = (time.perf_counter() - t0) / 3
    print(f"{'U_NZentries (fresh arrays)':26s} {dt * 1e3:8.2f} ms  {n / dt / 1e6:7.1f} Msets/s", flush=True)
