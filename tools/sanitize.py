"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gpvecchia_b200 as G
from gpvecchia_b200 import harness as H

CASES = ((600, 30, 2, "z"), (400, 12, 3, "z"), (300, 40, 2, "z"), (500, 9, 2, "zy"), (200, 5, 5, "z"),
         (300, 20, 2, "z"), (300, 25, 2, "z"), (300, 31, 2, "z"), (300, 40, 3, "z"), (300, 20, 3, "zy"))
# every case twice: plain path, and with the locality layer forced on (Morton-sorted replica + permuted set list)
for loc, (n, m, d, layout) in [(l, c) for l in ("0", "1") for c in CASES]:
    os.environ["GPV_LOCALITY"] = loc
    locs = H.make_locs(n, d, stream=1)
    if layout == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        nug_all = np.concatenate([np.full(n, 0.1), np.zeros(n)])
    else:
        NN = (H.rev(H.ordered_nn_gpu(locs, m)) if d <= 3 else H.ordered_nn_kdtree(locs, m)).astype(np.int64)
        locs2 = locs
        Cond, obs = H.layout_yz(NN, "z"), np.ones(n, dtype=bool)
        nug_all = np.full(n, 0.1)
    tau = nug_all[:n]
    z = H.make_data(n, stream=1)
    with G.UHandle(locs2, H.rev(NN), H.rev(Cond), obs=obs) as h:
        for ct, cp in (("matern", [1.0, 0.2, 1.5]), ("matern", [1.0, 0.2, 0.5]), ("matern", [1.0, 0.2, 2.5]),
                       ("matern", [1.0, 0.2, 0.8]), ("esqe", [0.7, 0.2, 0.4, 0.1])):
            h.U_NZentries(ct, cp, nug_all, tau)
            h.values_packed(ct, cp, nug_all, tau)
            h.values_csc(ct, cp, nug_all, tau)
            h.loglik_numerator(ct, cp, nug_all, tau, z, skip_rows=n if layout == "zy" else 0)
        h.u_sparsity(); h.csc_pattern()
        h.set_scalar_nugget(0.1)                       # resident scalar nugget: NULL vectors in the value calls
        h.values_packed("matern", [1.0, 0.2, 1.5], None, None)
        h.values_csc("matern", [1.0, 0.2, 1.5], None, None)
        if layout == "z":
            h.loglik_z("matern", [1.0, 0.2, 1.5], nug_all, tau, z)
# the chunked output pipelines: page-locked destination (copy stream, nugget upload staged with the chunks) and
# pageable destination of more than 8 MB (copy workers), locality layer off and on
import torch
for loc in ("0", "1"):
    os.environ["GPV_LOCALITY"] = loc
    n, m = 100000, 10
    locs = H.make_locs(n, 2, stream=3)
    NN = H.rev(H.ordered_nn_gpu(locs, m)).astype(np.int64)
    Cond = H.layout_yz(NN, "z")
    nug = np.full(n, 0.1)
    with G.UHandle(locs, H.rev(NN), H.rev(Cond), obs=np.ones(n, dtype=bool)) as h:
        total = h.packed_len + 2 * n
        pinned = torch.empty(total, dtype=torch.float64).pin_memory().numpy()
        h.values_packed("matern", [1.0, 0.02, 1.5], nug, nug, out=pinned)
        paged = np.empty(total)
        h.values_packed("matern", [1.0, 0.02, 1.5], nug, nug, out=paged)
        assert np.array_equal(pinned, paged)
        _, nnz, _ = h.csc_dims()
        h.values_csc("matern", [1.0, 0.02, 1.5], nug, nug, out=np.empty(nnz))
        h.U_NZentries("matern", [1.0, 0.02, 1.5], nug, nug)
G.MaternFun(np.linspace(0, 3, 100), [1.0, 0.3, 1.3])
G.EsqeFun(np.linspace(0, 3, 100), [1.0, 0.3, 0.5, 0.2])
print("sanitize workload done")
