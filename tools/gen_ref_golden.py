"""Generates tests/golden/ref_compiled.npz: outputs of the COMPILED REFERENCE (oracle/_ref: the unmodified
/root/reference/src/{U_NZentries,Matern,Esqe,dist,ic0}.cpp built by oracle/ref_build/Makefile) on the inputs of
tests/ref_cases.py, plus MaternFun / EsqeFun / ic0 / createUcpp / U_NZentries_mat outputs.  These are the
reference-run fixtures that pin the restatement (oracle/) and the CUDA path; /root/reference is only read here,
at generation time, never by a test.      Run (in the build container):  python tools/gen_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_native as R  # noqa: E402
import oracle as O  # noqa: E402
from ref_cases import cases  # noqa: E402

R.build()
assert R.has_lapack()
fix = {}
for name, c in cases().items():
    R.force_textbook(c["textbook"])
    r = R.U_NZentries(2, c["n"], c["locs"], c["revNNarray"], c["revCond"], c["nuggets"], c["nuggets_obsord"],
                      c["covType"], c["covparms"])
    R.force_textbook(False)
    for k in ("locs", "revNNarray", "revCond", "nuggets", "nuggets_obsord", "covparms"):
        fix[f"{name}/{k}"] = c[k]
    fix[f"{name}/covType"] = np.array(c["covType"])
    fix[f"{name}/textbook"] = np.array(c["textbook"])
    fix[f"{name}/Lentries"] = r["Lentries"]
    fix[f"{name}/Zentries"] = r["Zentries"]
    fix[f"{name}/nfail"] = np.array(r["nfail"])
    print(f"{name:28s} N={c['locs'].shape[0]:4d} p={c['revNNarray'].shape[1]:2d} nfail={r['nfail']}")

# covariance functions alone, on a distance matrix with zeros on and off the diagonal
rng = np.random.default_rng(11)
pts = rng.random((40, 2))
pts[17] = pts[4]
D = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
fix["cov/D"] = D
for nu in (0.5, 1.5, 2.5, 0.8, 1.3, 3.7):
    fix[f"cov/matern_{nu}"] = R.MaternFun(D, [1.7, 0.2, nu])
fix["cov/esqe"] = R.EsqeFun(D, [0.7, 0.25, 0.4, 0.6])

# U_NZentries_mat (covmodel given as a matrix, createU.R:149-151)
c = cases()["sgv_m10_nu25"]
N = c["locs"].shape[0]
Dm = np.sqrt(((c["locs"][:, None] - c["locs"][None]) ** 2).sum(-1))
covVals = O.MaternFun(Dm, [0.9, 0.3, 2.5]) + 0.05 * np.eye(N)
r = R.U_NZentries_mat(2, c["n"], c["locs"], c["revNNarray"], c["revCond"], c["nuggets"], c["nuggets_obsord"],
                      covVals, c["covparms"])
fix["mat/covVals"] = covVals
fix["mat/Lentries"] = r["Lentries"]
fix["mat/Zentries"] = r["Zentries"]

# ic0 / createUcpp: (a) the full lower-triangular pattern (ic0 = exact Cholesky, cf. tests/testthat/test-createL.r:43-45),
# (b) the pattern of an ordered-nearest-neighbour array, where ic0 may go indefinite and the reference then
# propagates NaN (sqrt of a negative pivot, src/ic0.cpp:55): that behaviour is part of the fixture.
def pattern(n, NNa):
    ptrs, inds = [0], []
    for i in range(n):
        cols = list(range(i + 1)) if NNa is None else sorted(int(v) - 1 for v in NNa[i] if v > 0)
        inds += cols
        ptrs.append(len(inds))
    return np.array(ptrs, dtype=np.float64), np.array(inds, dtype=np.float64)


for tag, n, m in (("ic0", 24, None), ("ic0nn", 60, 6)):
    locs = rng.random((n, 2))
    ptrs, inds = pattern(n, None if m is None else O.find_ordered_nn_brute(locs, m))
    cp = np.array([1.0, 0.3, 1.5])
    fix[f"{tag}/ptrs"], fix[f"{tag}/inds"], fix[f"{tag}/locs"], fix[f"{tag}/covparams"] = ptrs, inds, locs, cp
    fix[f"{tag}/createUcpp"] = R.createUcpp(ptrs, inds, locs, cp)
    Dfull = np.sqrt(((locs[:, None] - locs[None]) ** 2).sum(-1))
    rows = np.repeat(np.arange(n), np.diff(ptrs).astype(int))
    vals = R.MaternFun(Dfull, cp)[rows, inds.astype(int)]
    fix[f"{tag}/vals_in"] = vals
    fix[f"{tag}/ic0"] = R.ic0(ptrs, inds, vals)
    fix[f"{tag}/createUcppM"] = R.createUcppM(ptrs, inds, vals)
    print(tag, "finite:", bool(np.isfinite(fix[f"{tag}/ic0"]).all()))

out = os.path.join(ROOT, "tests", "golden", "ref_compiled.npz")
np.savez_compressed(out, **fix)
print("wrote", out, os.path.getsize(out), "bytes")
