"""End-to-end time of the R shim's own routines (r_shim/src/gpv_shim.c, run against the mock of R's C API of
tests/r_api_mock): `_GPvecchia_b200_U_values_csc` -- the call the drop-in createU() makes -- with an ordinary
(pageable, freshly allocated) result vector and with options(GPvecchia.b200.pinned_results = TRUE), and the
stateless `_GPvecchia_U_NZentries` with the reference's nine arguments (handle created and destroyed per call)."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_r_shim_mock import MockR, NA_INT
from gpvecchia_b200 import harness as H

n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 30
R = MockR()
locs_np = H.make_locs(n, 2, stream=2)
nn_np = H.ordered_nn_gpu(locs_np, m).astype(np.int32)          # reversed neighbour array, self last
nn_np[nn_np <= 0] = NA_INT
locs, nn, rc = R.real(locs_np), R.integer(nn_np), R.logical(np.zeros(nn_np.shape, np.int32))
tau = R.real(H.make_nuggets(n, stream=2))
cov, cp = R.string("matern"), R.real([1.0, H.default_range(n, 2), 1.5])
h = R.call("_GPvecchia_b200_create", locs, nn, rc, R.logical(np.ones(n, np.int32)))
R.call("_GPvecchia_b200_csc_pattern", h)


def timed(name, f, k=10, release=True):
    for _ in range(2):
        v = f()
        if release:
            R.L.mock_release(v)
    t0 = time.perf_counter()
    for _ in range(k):
        v = f()
        if release:
            R.L.mock_release(v)          # the garbage collector: a pooled block goes back to the shim
    dt = (time.perf_counter() - t0) / k
    print(f"{name:64s} {dt * 1e3:8.2f} ms  {n / dt / 1e6:7.1f} Msets/s", flush=True)


timed("_GPvecchia_b200_U_values_csc, ordinary R vector", lambda: R.call("_GPvecchia_b200_U_values_csc", h, cov, cp, tau, tau))
timed("_GPvecchia_b200_U_values, ordinary R vector", lambda: R.call("_GPvecchia_b200_U_values", h, cov, cp, tau, tau))
R.option("GPvecchia.b200.pinned_results", R.logical([1]))
timed("_GPvecchia_b200_U_values_csc, pinned_results = TRUE", lambda: R.call("_GPvecchia_b200_U_values_csc", h, cov, cp, tau, tau))
timed("_GPvecchia_b200_U_values, pinned_results = TRUE", lambda: R.call("_GPvecchia_b200_U_values", h, cov, cp, tau, tau))
R.option("GPvecchia.b200.pinned_results", None)
timed("_GPvecchia_U_NZentries (9 arguments, handle per call)",
      lambda: R.call("_GPvecchia_U_NZentries", R.integer([1]), R.real([float(n)]), locs, nn, rc, tau, tau, cov, cp), k=5, release=False)
R.L.mock_run_finalizer(h)
