"""The R shim, executed.

r_shim/src/gpv_shim.c is the `.Call` side of the drop-in boundary: what a GPvecchia maintainer compiles into the
package.  There is no R in this image, so the file is built UNMODIFIED against an executable stand-in for the R C
API (tests/r_api_mock/mock_runtime.c: vectors, attributes, coercion that returns its argument when the type
matches, NA_integer_, Rf_error as a longjmp, external pointers with finalizers, the routine table of
R_registerRoutines) and its routines are called BY THEIR REGISTERED NAMES, as `.Call("_GPvecchia_U_NZentries", ...)`
resolves them (src/RcppExports.cpp:155-172).

CPU tests: registration names and arities, argument errors as R errors, ic0 through the shim (host code).
GPU tests: every routine against the oracle / the ctypes front end on the same inputs, including NA_integer_
neighbour ids, the caller's matrices left untouched, stale handles, finalizers and the scalar-nugget route.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import gpvecchia_b200 as G
import oracle as O
from gpvecchia_b200 import harness as H

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK = os.path.join(HERE, "r_api_mock")
LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 10, 13, 14, 16, 19
NA_INT = np.iinfo(np.int32).min
VAL_TOL = 1e-10


class RError(Exception):
    pass


class MockR:
    """The few things an R session does around `.Call`: build vectors, call a registered routine, read results."""

    def __init__(self):
        if shutil.which("gcc") is None or shutil.which("make") is None:
            pytest.skip("no gcc / make")
        r = subprocess.run(["make", "-C", MOCK], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        L = C.CDLL(os.path.join(MOCK, "_build", "libgpv_shim_mock.so"))
        P = C.c_void_p
        for name, res, args in (
                ("mock_nil", P, []), ("mock_alloc", P, [C.c_int, C.c_int64]), ("mock_set_dim", None, [P, C.c_int, C.c_int]),
                ("mock_string", P, [C.c_char_p]), ("mock_data", P, [P]), ("mock_length", C.c_int64, [P]),
                ("mock_type", C.c_int, [P]), ("mock_elt", P, [P, C.c_int64]), ("mock_chars", C.c_char_p, [P]),
                ("mock_names", P, [P]), ("mock_extptr_addr", P, [P]), ("mock_call", P, [C.c_char_p, C.c_int, C.POINTER(P)]),
                ("mock_last_error", C.c_char_p, []), ("mock_last_warning", C.c_char_p, []), ("mock_warning_count", C.c_int, []),
                ("mock_protect_depth", C.c_int, []), ("mock_routine_arity", C.c_int, [C.c_char_p]),
                ("mock_routine_count", C.c_int, []), ("mock_routine_name", C.c_char_p, [C.c_int]),
                ("mock_set_option", None, [C.c_char_p, P]), ("mock_run_finalizer", None, [P]), ("mock_null_extptr", None, [P]),
                ("mock_release", None, [P]), ("Rf_nrows", C.c_int, [P]), ("Rf_ncols", C.c_int, [P]), ("Rf_isMatrix", C.c_int, [P]),
                ("R_init_GPvecchiaB200", None, [P])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.L = L
        L.R_init_GPvecchiaB200(None)                 # what R does when it loads the package's shared object
        self.nil = L.mock_nil()

    def _vec(self, sxptype, a, dtype):
        a = np.asarray(a)
        flat = np.ascontiguousarray(a.T if a.ndim == 2 else a, dtype=dtype).ravel()     # R matrices are column-major
        s = self.L.mock_alloc(sxptype, flat.size)
        if flat.size:
            C.memmove(self.L.mock_data(s), flat.ctypes.data, flat.nbytes)
        if a.ndim == 2:
            self.L.mock_set_dim(s, a.shape[0], a.shape[1])
        return s

    def real(self, a):
        return self._vec(REALSXP, a, np.float64)

    def integer(self, a):
        return self._vec(INTSXP, a, np.int32)

    def logical(self, a):
        return self._vec(LGLSXP, a, np.int32)

    def string(self, s):
        return self.L.mock_string(s.encode())

    def option(self, name, value):
        self.L.mock_set_option(name.encode(), self.nil if value is None else value)

    def call(self, name, *args):
        arr = (C.c_void_p * max(len(args), 1))(*args)
        depth = self.L.mock_protect_depth()
        r = self.L.mock_call(name.encode(), len(args), arr)
        assert self.L.mock_protect_depth() == depth, "PROTECT / UNPROTECT out of balance in " + name
        if not r:
            raise RError(self.L.mock_last_error().decode())
        return r

    def numpy(self, s):
        t, n = self.L.mock_type(s), self.L.mock_length(s)
        dt = {REALSXP: np.float64, INTSXP: np.int32, LGLSXP: np.int32}[t]
        a = np.frombuffer(C.string_at(self.L.mock_data(s), n * np.dtype(dt).itemsize), dtype=dt).copy()
        if self.L.Rf_isMatrix(s):
            a = a.reshape(self.L.Rf_ncols(s), self.L.Rf_nrows(s)).T
        return a

    def view(self, s, dtype):
        """The object's own storage (no copy): to see whether a routine wrote through its argument."""
        n = self.L.mock_length(s)
        return np.ctypeslib.as_array(C.cast(self.L.mock_data(s), C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))

    def list_(self, s):
        names = self.L.mock_names(s)
        return {self.L.mock_chars(self.L.mock_elt(names, i)).decode(): self.L.mock_elt(s, i)
                for i in range(self.L.mock_length(s))}


@pytest.fixture(scope="module")
def R():
    return MockR()


# ---- CPU: registration, argument errors, host-side routines ------------------------------------------
def test_registered_names_and_arities_are_the_reference_ones(R):
    # src/RcppExports.cpp:155-172 (R_CallMethodDef CallEntries[]): same names, same number of arguments
    for name, arity in (("_GPvecchia_U_NZentries", 9), ("_GPvecchia_U_NZentries_mat", 9), ("_GPvecchia_MaternFun", 2),
                        ("_GPvecchia_EsqeFun", 2), ("_GPvecchia_ic0", 3), ("_GPvecchia_createUcppM", 3),
                        ("_GPvecchia_createUcpp", 4)):
        assert R.L.mock_routine_arity(name.encode()) == arity, name
    names = [R.L.mock_routine_name(i).decode() for i in range(R.L.mock_routine_count())]
    assert len(names) == len(set(names)) and all(n.startswith("_GPvecchia_") for n in names)
    with pytest.raises(RError, match="not available"):
        R.call("_GPvecchia_no_such_routine", R.nil)
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("_GPvecchia_ic0", R.nil, R.nil)


def test_ic0_through_the_shim_overwrites_its_argument_like_the_reference(R):
    # src/ic0.cpp:43-63: NumericVector arguments wrap the R vectors, vals is overwritten and returned
    rng = np.random.default_rng(5)
    N = 40
    A = rng.standard_normal((N, N))
    S = A @ A.T + N * np.eye(N)
    ptrs, inds, vals = [0], [], []
    for i in range(N):
        cols = [j for j in range(i + 1) if j == i or rng.random() < 0.4]
        inds += cols
        vals += [S[i, j] for j in cols]
        ptrs.append(len(inds))
    want = O.ic0(np.array(ptrs, float), np.array(inds, float), np.array(vals, float))
    v = R.real(vals)
    out = R.call("_GPvecchia_ic0", R.real(np.array(ptrs, float)), R.real(np.array(inds, float)), v)
    assert out == v                                                   # the same R object comes back
    assert np.array_equal(R.numpy(v), want)
    # integer ptrs / inds (as.integer on the R side) are coerced, a copy of integer vals is returned instead
    vi = R.integer(np.arange(len(inds)) + 1)
    out2 = R.call("_GPvecchia_createUcppM", R.integer(ptrs), R.integer(inds), vi)
    assert out2 != vi and np.array_equal(R.view(vi, np.int32), np.arange(len(inds)) + 1)
    with pytest.raises(RError, match="differ in length"):
        R.call("_GPvecchia_ic0", R.real(np.array(ptrs, float)), R.real(np.array(inds, float)), R.real(vals[:-1]))


def test_argument_errors_are_r_errors_not_crashes(R):
    locs = R.real(np.zeros((4, 2)))
    with pytest.raises(RError, match="locs must be a numeric matrix"):
        R.call("_GPvecchia_U_NZentries", R.integer([1]), R.real([4.0]), R.real(np.zeros(8)), R.integer(np.ones((4, 2))),
               R.logical(np.zeros((4, 2))), R.real(np.ones(4)), R.real(np.ones(4)), R.string("matern"), R.real([1, 1, 1.5]))
    with pytest.raises(RError, match="revNNarray must be a matrix"):
        R.call("_GPvecchia_U_NZentries", R.integer([1]), R.real([4.0]), locs, R.integer(np.ones(8)),
               R.logical(np.zeros((4, 2))), R.real(np.ones(4)), R.real(np.ones(4)), R.string("matern"), R.real([1, 1, 1.5]))
    with pytest.raises(RError, match="EsqeFun"):
        R.call("_GPvecchia_EsqeFun", R.real(np.ones((2, 2))), R.real([1.0, 2.0]))
    if G.lib.gpv_device_count() == 0:
        # no GPU: the product fails loudly, as an R error that carries the library's message
        with pytest.raises(RError, match="gpvecchia_b200"):
            R.call("_GPvecchia_b200_create", locs, R.integer(np.ones((4, 2))), R.logical(np.zeros((4, 2))),
                   R.logical(np.ones(4)))


# ---- GPU ------------------------------------------------------------------------------------------------
def _problem(n, m, d, cond_yz, stream):
    locs = H.make_locs(n, d, stream=stream)
    NN = H.ordered_nn_kdtree(locs, m)
    Cond = O.whichCondOnLatent(NN) if cond_yz == "SGV" else H.layout_yz(NN, cond_yz)
    return H.make_vecchia_approx(locs, NN, Cond, np.ones(n, dtype=bool), cond_yz)


def _r_arrays(R, va):
    """locsord, revNNarray (integer, NA_integer_ where R has NA) and revCond (logical) as R would hold them."""
    prep = va["U_prep"]
    nn = np.asarray(prep["revNNarray"]).astype(np.int32)
    nn[nn <= 0] = NA_INT
    rc = np.asarray(prep["revCond"]).astype(np.int32)
    rc[np.asarray(prep["revCond"]) < 0] = NA_INT
    return R.real(va["locsord"]), R.integer(nn), R.logical(rc), nn, rc


def _rowscaled_err(got, ref):
    scale = np.abs(ref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return float((np.abs(got - ref) / scale).max())


@pytest.mark.gpu
@pytest.mark.parametrize("covType,cp", [("matern", [1.3, 0.05, 1.5]), ("matern", [1.0, 0.05, 0.8]),
                                        ("esqe", [0.7, 0.05, 0.4, 0.11])])
def test_U_NZentries_by_its_reference_name_against_the_oracle(R, covType, cp):
    n, m = 1500, 12
    va = _problem(n, m, 2, "SGV", stream=31)
    prep = va["U_prep"]
    locs, nn, rc, nn_np, rc_np = _r_arrays(R, va)
    nug = H.make_nuggets(n, stream=31)
    out = R.call("_GPvecchia_U_NZentries", R.integer([4]), R.real([float(n)]), locs, nn, rc, R.real(nug), R.real(nug),
                 R.string(covType), R.real(cp))
    parts = R.list_(out)
    assert list(parts) == ["Lentries", "Zentries"]                       # List::create(Named(..)) of U_NZentries.cpp:117
    Lg, Zg = R.numpy(parts["Lentries"]), R.numpy(parts["Zentries"])
    assert Lg.shape == (n, m + 1) and Zg.shape == (2 * n, 1)
    rcd = np.asarray(prep["revCond"]).astype(np.float64)
    ref = O.U_NZentries(O.max_threads(), n, va["locsord"], prep["revNNarray"], rcd, nug, nug, covType, np.asarray(cp, float))
    assert np.array_equal(Lg == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(Lg, ref["Lentries"]) < VAL_TOL
    assert np.allclose(Zg.ravel(), ref["Zentries"], rtol=2e-16, atol=0)
    # the caller's matrices are untouched: is.na(revNNarray) still holds afterwards (createU.R:158 needs it)
    assert np.array_equal(R.view(nn, np.int32), nn_np.T.ravel())
    assert np.array_equal(R.view(rc, np.int32), rc_np.T.ravel())
    # a double revNNarray (what t(apply(., 1, rev)) of a numeric matrix gives) is coerced, NA and all
    nnd = nn_np.astype(np.float64)
    nnd[nn_np == NA_INT] = np.nan
    out2 = R.call("_GPvecchia_U_NZentries", R.integer([1]), R.real([float(n)]), locs, R.real(nnd), rc, R.real(nug),
                  R.real(nug), R.string(covType), R.real(cp))
    assert np.array_equal(R.numpy(R.list_(out2)["Lentries"]), Lg)
    with pytest.raises(RError, match="gpvecchia_b200"):
        R.call("_GPvecchia_U_NZentries", R.integer([1]), R.real([float(n)]), locs, nn, rc, R.real(nug), R.real(nug),
               R.string("no_such_covariance"), R.real(cp))


@pytest.mark.gpu
def test_handle_routines_build_the_same_U_as_the_ctypes_front_end(R):
    import scipy.sparse as sp
    n, m = 2500, 15
    va = _problem(n, m, 2, "SGV", stream=32)
    prep = va["U_prep"]
    cp = [1.1, 0.04, 1.5]
    locs, nn, rc, nn_np, _ = _r_arrays(R, va)
    h = R.call("_GPvecchia_b200_create", locs, nn, rc, R.logical(np.asarray(va["obs"]).astype(np.int32)))
    assert np.array_equal(R.view(nn, np.int32), nn_np.T.ravel())          # not written through (Rf_coerceVector returned nn itself)
    size = prep["size"]
    want = G.createU(va, cp, 0.3)["U"].tocsc()
    want.sort_indices()
    pat = R.call("_GPvecchia_b200_csc_pattern", h)
    p_, i_ = R.numpy(R.L.mock_elt(pat, 0)), R.numpy(R.L.mock_elt(pat, 1))
    # per-location vectors ...
    nug_all = np.full(n, 0.3)
    x = R.numpy(R.call("_GPvecchia_b200_U_values_csc", h, R.string("matern"), R.real(cp), R.real(nug_all), R.real(nug_all)))
    U = sp.csc_matrix((x, i_, p_), shape=(size, size))
    U.sort_indices()
    assert np.array_equal(U.indptr, want.indptr) and np.array_equal(U.indices, want.indices)
    assert np.array_equal(U.data, want.data)
    # ... or the scalar nugget built on the device and NULL vectors: the same bits
    R.call("_GPvecchia_b200_set_scalar_nugget", h, R.real([0.3]))
    x2 = R.numpy(R.call("_GPvecchia_b200_U_values_csc", h, R.string("matern"), R.real(cp), R.nil, R.nil))
    assert np.array_equal(x2, x)
    # triplet route: allLentries + the arrays of U_sparsity.R
    al = R.numpy(R.call("_GPvecchia_b200_U_values", h, R.string("matern"), R.real(cp), R.nil, R.nil))
    spars = R.call("_GPvecchia_b200_U_sparsity", h)
    ci, rp = R.numpy(R.L.mock_elt(spars, 0)), R.numpy(R.L.mock_elt(spars, 1))
    assert np.array_equal(ci, prep["colindices"]) and np.array_equal(rp, prep["rowpointers"])
    U3 = sp.coo_matrix((al, (ci - 1, rp - 1)), shape=(size, size)).tocsc()
    U3.sort_indices()
    assert np.array_equal(U3.data, want.data)
    # likelihood: whole value for the pure `z` layout is refused here (SGV), the numerator works
    z = H.make_data(n, stream=32)
    num = R.numpy(R.call("_GPvecchia_b200_loglik_numerator", h, R.string("matern"), R.real(cp), R.nil, R.nil,
                         R.real(z[np.asarray(va["ord_z"]) - 1]), R.real([0.0])))
    q, l, nf = G.vecchia_loglik_numerator(z, va, cp, 0.3)
    assert num[2] == nf == 0 and abs(num[0] - q) <= 1e-12 * abs(q) and abs(num[1] - l) <= 1e-12 * abs(l)
    with pytest.raises(RError, match="gpvecchia_b200"):
        R.call("_GPvecchia_b200_loglik_z", h, R.string("matern"), R.real(cp), R.nil, R.nil, R.nil)
    # a failing Cholesky is an R warning, not an error (the reference prints to Rcerr and zeroes the row)
    w0 = R.L.mock_warning_count()
    R.call("_GPvecchia_b200_U_values", h, R.string("matern"), R.real([1.0, 0.04, 1.5]), R.real(np.full(n, -5.0)), R.real(nug_all))
    assert R.L.mock_warning_count() == w0 + 1 and b"Cholesky decomposition failed" in R.L.mock_last_warning()
    # what readRDS() gives back is a NULL external pointer: an R error that says so, not a crash
    addr = R.L.mock_extptr_addr(h)
    assert addr
    R.L.mock_run_finalizer(h)                                             # the GC's job: frees the device handle once
    assert not R.L.mock_extptr_addr(h)
    R.L.mock_run_finalizer(h)                                             # idempotent
    with pytest.raises(RError, match="stale device handle"):
        R.call("_GPvecchia_b200_csc_pattern", h)


@pytest.mark.gpu
def test_result_vectors_in_page_locked_memory_through_a_custom_r_allocator(R):
    # options(GPvecchia.b200.pinned_results = TRUE): the storage of the result vector comes from gpv_host_alloc through
    # Rf_allocVector3, goes back to the shim's pool when the vector is collected, and is reused by the next call
    n, m = 100000, 12
    va = _problem(n, m, 2, "z", stream=35)
    cp = [1.0, 0.01, 1.5]
    locs, nn, rc, _, _ = _r_arrays(R, va)
    h = R.call("_GPvecchia_b200_create", locs, nn, rc, R.logical(np.ones(n, np.int32)))
    tau = np.full(n, 0.2)
    args = (h, R.string("matern"), R.real(cp), R.real(tau), R.real(tau))
    want = R.numpy(R.call("_GPvecchia_b200_U_values", *args))
    want_x = R.numpy(R.call("_GPvecchia_b200_U_values_csc", *args))
    R.option("GPvecchia.b200.pinned_results", R.logical([1]))
    try:
        v1 = R.call("_GPvecchia_b200_U_values", *args)
        p1 = R.L.mock_data(v1)
        assert np.array_equal(R.numpy(v1), want)
        v2 = R.call("_GPvecchia_b200_U_values", *args)                    # v1 still alive: a second block
        assert R.L.mock_data(v2) != p1 and np.array_equal(R.numpy(v2), want)
        R.L.mock_release(v1)                                              # the garbage collector frees v1 ...
        v3 = R.call("_GPvecchia_b200_U_values", *args)                    # ... and its block serves the next call
        assert R.L.mock_data(v3) == p1 and np.array_equal(R.numpy(v3), want)
        x = R.call("_GPvecchia_b200_U_values_csc", *args)
        assert np.array_equal(R.numpy(x), want_x)
        for v in (v2, v3, x):
            R.L.mock_release(v)
    finally:
        R.option("GPvecchia.b200.pinned_results", None)
    R.L.mock_run_finalizer(h)


@pytest.mark.gpu
def test_whole_likelihood_and_covariance_functions_through_the_shim(R):
    n, m = 3000, 10
    va = _problem(n, m, 2, "z", stream=33)
    cp = [0.9, 0.03, 0.8]
    locs, nn, rc, _, _ = _r_arrays(R, va)
    h = R.call("_GPvecchia_b200_create", locs, nn, rc, R.logical(np.ones(n, np.int32)))
    z = H.make_data(n, stream=33)
    zord = z[np.asarray(va["ord_z"]) - 1]
    tau = np.full(n, 0.2)
    r = R.numpy(R.call("_GPvecchia_b200_loglik_z", h, R.string("matern"), R.real(cp), R.real(tau), R.real(tau), R.real(zord)))
    want = O.vecchia_likelihood_U(z, O.createU(va, cp, 0.2))
    assert r[5] == 0 and abs(r[0] - want) <= 1e-8 * abs(want)
    # estimation loop: scalar nugget on the device, z resident -> NULL, NULL, NULL
    R.call("_GPvecchia_b200_set_scalar_nugget", h, R.real([0.2]))
    r2 = R.numpy(R.call("_GPvecchia_b200_loglik_z", h, R.string("matern"), R.real(cp), R.nil, R.nil, R.nil))
    assert r2[0] == r[0]
    R.L.mock_run_finalizer(h)
    # MaternFun / EsqeFun keep the shape of distmat (Matern.cpp:24, Esqe.cpp:17)
    D = np.abs(np.random.default_rng(3).standard_normal((7, 5)))
    D[0, 0] = 0.0
    for nu in (0.5, 1.5, 2.5, 1.3):
        got = R.numpy(R.call("_GPvecchia_MaternFun", R.real(D), R.real([1.4, 0.7, nu])))
        assert got.shape == D.shape and np.allclose(got, O.MaternFun(D, np.array([1.4, 0.7, nu])), rtol=1e-13, atol=1e-15)
    got = R.numpy(R.call("_GPvecchia_EsqeFun", R.real(D), R.real([1.4, 0.7, 0.3, 0.2])))
    assert np.allclose(got, O.EsqeFun(D, np.array([1.4, 0.7, 0.3, 0.2])), rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
def test_device_options_select_one_or_several_gpus(R):
    # options(GPvecchia.b200.devices = c(0, ...)): one R process, a worker thread per listed device inside the library
    ndev = G.lib.gpv_device_count()
    n, m = 4000, 12
    va = _problem(n, m, 2, "z", stream=34)
    cp = [1.0, 0.03, 1.5]
    locs, nn, rc, _, _ = _r_arrays(R, va)
    obs = R.logical(np.ones(n, np.int32))
    tau = np.full(n, 0.25)
    h1 = R.call("_GPvecchia_b200_create", locs, nn, rc, obs)
    x1 = R.numpy(R.call("_GPvecchia_b200_U_values", h1, R.string("matern"), R.real(cp), R.real(tau), R.real(tau)))
    devs = list(range(ndev)) if ndev >= 2 else [0, 0]
    R.option("GPvecchia.b200.devices", R.integer(devs))
    try:
        hm = R.call("_GPvecchia_b200_create", locs, nn, rc, obs)
        xm = R.numpy(R.call("_GPvecchia_b200_U_values", hm, R.string("matern"), R.real(cp), R.real(tau), R.real(tau)))
        assert np.array_equal(xm, x1)
        pat1, patm = R.call("_GPvecchia_b200_csc_pattern", h1), R.call("_GPvecchia_b200_csc_pattern", hm)
        for k in (0, 1):
            assert np.array_equal(R.numpy(R.L.mock_elt(pat1, k)), R.numpy(R.L.mock_elt(patm, k)))
        z = H.make_data(n, stream=34)
        zord = z[np.asarray(va["ord_z"]) - 1]
        a = R.numpy(R.call("_GPvecchia_b200_loglik_z", h1, R.string("matern"), R.real(cp), R.real(tau), R.real(tau), R.real(zord)))
        b = R.numpy(R.call("_GPvecchia_b200_loglik_z", hm, R.string("matern"), R.real(cp), R.real(tau), R.real(tau), R.real(zord)))
        assert abs(a[0] - b[0]) <= 1e-12 * abs(a[0])
        with pytest.raises(RError, match="single-device handle only"):
            R.call("_GPvecchia_b200_set_scalar_nugget", hm, R.real([0.25]))
        R.L.mock_run_finalizer(hm)
    finally:
        R.option("GPvecchia.b200.devices", None)
    R.L.mock_run_finalizer(h1)
