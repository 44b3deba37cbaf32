"""Pins the restatement (oracle/) to the REFERENCE ITSELF: tests/golden/ref_compiled.npz holds outputs of the
unmodified /root/reference/src/{U_NZentries,Matern,Esqe,dist,ic0}.cpp compiled by oracle/ref_build (generator:
tools/gen_ref_golden.py; inputs: tests/ref_cases.py).  The restatement must reproduce them BIT FOR BIT -- same
LAPACK (the OpenBLAS inside scipy), same libm, same operation order -- and, where the compiled reference can
be loaded (oracle/_ref/libgpvecchia_ref.so: built in the container that has /root/reference, prebuilt on the GPU
box), it is also run live beside the restatement on fresh inputs.  /root/reference is never read by a test."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import ref_native as R
from ref_cases import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_compiled.npz")
FIX = np.load(GOLD)
NAMES = sorted({k.split("/")[0] for k in FIX.files} - {"cov", "mat", "ic0", "ic0nn"})
live = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(R.__file__), "_ref", "libgpvecchia_ref.so"))
                          and not os.path.isdir("/root/reference"),
                          reason="compiled reference (oracle/_ref) neither prebuilt nor buildable here")


def _case(name):
    g = lambda k: FIX[f"{name}/{k}"]                                  # noqa: E731
    return dict(n=g("nuggets_obsord").size, locs=g("locs"), revNNarray=g("revNNarray"), revCond=g("revCond"),
                nuggets=g("nuggets"), nuggets_obsord=g("nuggets_obsord"), covType=str(g("covType")),
                covparms=g("covparms"), textbook=bool(g("textbook")), Lentries=g("Lentries"),
                Zentries=g("Zentries"), nfail=int(g("nfail")))


def test_fixture_inputs_are_the_committed_cases():
    cs = cases()
    assert sorted(cs) == NAMES
    for name, c in cs.items():
        f = _case(name)
        for k in ("locs", "revNNarray", "nuggets", "nuggets_obsord", "covparms"):
            assert np.array_equal(np.asarray(c[k]), f[k]), (name, k)
        assert np.array_equal(np.isnan(c["revCond"]), np.isnan(f["revCond"]))
        assert np.array_equal(np.nan_to_num(c["revCond"]), np.nan_to_num(f["revCond"]))


@pytest.mark.parametrize("name", NAMES)
def test_restatement_reproduces_the_reference_run_bit_for_bit(name):
    assert O.has_lapack()
    f = _case(name)
    r = O.U_NZentries(2, f["n"], f["locs"], f["revNNarray"], f["revCond"], f["nuggets"], f["nuggets_obsord"],
                      f["covType"], f["covparms"], mode=1 if f["textbook"] else 0)
    assert r["nfail"] == f["nfail"]
    assert np.array_equal(r["Lentries"], f["Lentries"], equal_nan=True), \
        float(np.nanmax(np.abs(r["Lentries"] - f["Lentries"])))
    assert np.array_equal(r["Zentries"], f["Zentries"], equal_nan=True)
    # the pattern createU.R:158 relies on: values in the first n0 slots, zeros after
    n0 = (f["revNNarray"] != 0).sum(axis=1)
    failed = (f["Lentries"] == 0).all(axis=1)
    cols = np.arange(f["Lentries"].shape[1])[None, :]
    if np.isfinite(f["nuggets"]).all():         # an Inf nugget gives legitimate exact zeros inside a row
        assert np.all((f["Lentries"] != 0) == ((cols < n0[:, None]) & ~failed[:, None]))


def test_covariance_restatements_reproduce_the_reference_run():
    D = FIX["cov/D"]
    for nu in (0.5, 1.5, 2.5, 0.8, 1.3, 3.7):
        assert np.array_equal(O.MaternFun(D, [1.7, 0.2, nu]), FIX[f"cov/matern_{nu}"]), nu
    assert np.array_equal(O.EsqeFun(D, [0.7, 0.25, 0.4, 0.6]), FIX["cov/esqe"])
    # the reference's own known answer (tests/testthat/test-MaternFun.r:32-41) holds for the reference-run values
    s = D / 0.2
    naive15 = 1.7 * (1 + np.sqrt(3) * s) * np.exp(-np.sqrt(3) * s)
    assert np.abs(naive15 - FIX["cov/matern_1.5"]).sum() < 1e-10


def test_mat_and_ic0_restatements_reproduce_the_reference_run():
    c = _case("sgv_m10_nu25")
    r = O.U_NZentries_mat(c["n"], c["revNNarray"], FIX["mat/covVals"], c["nuggets_obsord"])
    scale = np.abs(FIX["mat/Lentries"]).max(axis=1, keepdims=True)                   # numpy chol, not LAPACK potrf('U'):
    assert (np.abs(r["Lentries"] - FIX["mat/Lentries"]) / scale).max() < 1e-12         # rounding-level, not bit-equal
    assert np.array_equal(r["Lentries"] == 0, FIX["mat/Lentries"] == 0)
    assert np.array_equal(np.asarray(r["Zentries"]).ravel(), FIX["mat/Zentries"].ravel())
    assert np.isfinite(FIX["ic0/ic0"]).all() and not np.isfinite(FIX["ic0nn/ic0"]).all()
    for tag in ("ic0", "ic0nn"):
        ptrs, inds = FIX[f"{tag}/ptrs"], FIX[f"{tag}/inds"]
        assert np.array_equal(O.ic0(ptrs, inds, FIX[f"{tag}/vals_in"]), FIX[f"{tag}/ic0"], equal_nan=True)
        assert np.array_equal(O.createUcppM(ptrs, inds, FIX[f"{tag}/vals_in"]), FIX[f"{tag}/createUcppM"], equal_nan=True)
        assert np.array_equal(O.createUcpp(ptrs, inds, FIX[f"{tag}/locs"], FIX[f"{tag}/covparams"]),
                              FIX[f"{tag}/createUcpp"], equal_nan=True)
    # full pattern: L L^T = Sigma (the reference's known answer, test-createL.r:43-45, on its own output)
    n = FIX["ic0/locs"].shape[0]
    L = np.zeros((n, n))
    rows = np.repeat(np.arange(n), np.diff(FIX["ic0/ptrs"]).astype(int))
    L[rows, FIX["ic0/inds"].astype(int)] = FIX["ic0/ic0"]
    D = np.sqrt(((FIX["ic0/locs"][:, None] - FIX["ic0/locs"][None]) ** 2).sum(-1))
    assert np.abs(L @ L.T - O.MaternFun(D, FIX["ic0/covparams"])).max() < 1e-10


@live
def test_compiled_reference_reproduces_its_own_fixture():
    assert R.has_lapack() and R.lib().gpv_ref_openmp() == 1
    for name in NAMES:
        f = _case(name)
        R.force_textbook(f["textbook"])
        try:
            r = R.U_NZentries(3, f["n"], f["locs"], f["revNNarray"], f["revCond"], f["nuggets"], f["nuggets_obsord"],
                              f["covType"], f["covparms"])
        finally:
            R.force_textbook(False)
        assert np.array_equal(r["Lentries"], f["Lentries"], equal_nan=True) and r["nfail"] == f["nfail"], name


@live
@pytest.mark.parametrize("seed", range(6))
def test_compiled_reference_and_restatement_agree_on_fresh_inputs(seed):
    """Random shapes / layouts / covariances: restatement == compiled reference, bit for bit, both LAPACK backed."""
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.integers(1, 4))
    n = int(rng.integers(60, 400))
    m = int(rng.integers(1, 45))
    cond = ["z", "y", "SGV", "zy"][seed % 4]
    locs = rng.random((n, d))
    va = O.vecchia_specify(locs, min(m, n - 1), cond_yz=cond)
    prep = va["U_prep"]
    rc = prep["revCond"].astype(np.float64)
    rc[prep["revCond"] < 0] = np.nan
    N = va["locsord"].shape[0]
    nug = 0.05 + 0.1 * rng.random(N)
    if cond == "zy":
        nug[n:] = 0.0
    nobs = int(np.asarray(va["obs"]).sum())
    for ct, cp in [("matern", [1.2, 0.3, 0.5]), ("matern", [1.2, 0.3, 1.5]), ("matern", [0.8, 0.3, 2.5]),
                   ("matern", [1.0, 0.3, float(rng.uniform(0.2, 4.0))]), ("esqe", [0.7, 0.3, 0.4, 0.5])]:
        a = O.U_NZentries(4, nobs, va["locsord"], prep["revNNarray"], rc, nug, nug[:nobs], ct, np.array(cp))
        b = R.U_NZentries(4, nobs, va["locsord"], prep["revNNarray"], rc, nug, nug[:nobs], ct, np.array(cp))
        assert np.array_equal(a["Lentries"], b["Lentries"], equal_nan=True), (ct, cp)
        assert np.array_equal(a["Zentries"], b["Zentries"]) and a["nfail"] == b["nfail"]


@live
def test_compiled_reference_ic0_reproduces_its_fixture():
    # (an unknown covType is not run through the compiled reference: U_NZentries.cpp:27-29 only prints, the empty
    # covmat then makes solve() throw a logic_error inside the OpenMP region, which is not the runtime_error :64
    # catches, and the process aborts.  The C ABI returns GPV_ERR_UNSUPPORTED instead; tests/test_capi_cpu.py.)
    for tag in ("ic0", "ic0nn"):
        ptrs, inds = FIX[f"{tag}/ptrs"], FIX[f"{tag}/inds"]
        assert np.array_equal(R.ic0(ptrs, inds, FIX[f"{tag}/vals_in"]), FIX[f"{tag}/ic0"], equal_nan=True)
        assert np.array_equal(R.createUcpp(ptrs, inds, FIX[f"{tag}/locs"], FIX[f"{tag}/covparams"]),
                              FIX[f"{tag}/createUcpp"], equal_nan=True)
