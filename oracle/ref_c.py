"""ctypes front end of oracle/liboracle.so (the C++ restatement of src/U_NZentries.cpp,
src/dist.cpp, src/Matern.cpp, src/Esqe.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _find_lapack():
    try:
        import scipy
        pats = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                      "libscipy_openblas*.so"))
        if pats:
            return os.path.abspath(pats[0])
    except Exception:
        pass
    return None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    L.gpv_oracle_bind_lapack.argtypes = [C.c_char_p]
    L.gpv_oracle_bind_lapack.restype = C.c_int
    L.gpv_oracle_has_lapack.restype = C.c_int
    L.gpv_oracle_max_threads.restype = C.c_int
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.gpv_oracle_MaternFun_flat.argtypes = [dp, C.c_long, dp, dp]
    L.gpv_oracle_EsqeFun_flat.argtypes = [dp, C.c_long, dp, dp]
    L.gpv_oracle_U_NZentries.argtypes = [C.c_int, C.c_long, C.c_long, C.c_int, C.c_int,
                                         dp, ip, dp, dp, dp, C.c_char_p, dp, dp, dp, C.c_int,
                                         C.POINTER(C.c_long)]
    L.gpv_oracle_U_NZentries.restype = C.c_int
    L.gpv_oracle_U_NZentries_rows.argtypes = [C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_long,
                                              C.c_long, dp, ip, dp, dp, dp, C.c_char_p, dp, dp, dp,
                                              C.c_int, C.POINTER(C.c_long)]
    L.gpv_oracle_U_NZentries_rows.restype = C.c_int
    L.gpv_oracle_block_cond_proxy.argtypes = [C.c_long, C.c_long, C.c_int, C.c_int, dp, ip, dp, dp,
                                              C.c_char_p, dp]
    L.gpv_oracle_block_cond_proxy.restype = C.c_double
    L.gpv_oracle_ic0.argtypes = [C.c_long, dp, dp, dp]
    L.gpv_oracle_ic0.restype = C.c_long
    L.gpv_oracle_createUcpp.argtypes = [C.c_long, C.c_int, dp, dp, dp, dp, C.c_int, dp]
    L.gpv_oracle_createUcpp.restype = C.c_long
    lp = _find_lapack()
    if lp is not None:
        L.gpv_oracle_bind_lapack(lp.encode())
    _LIB = L
    return L


def has_lapack():
    return bool(lib().gpv_oracle_has_lapack())


def max_threads():
    return int(lib().gpv_oracle_max_threads())


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def MaternFun(distmat, covparms):
    """src/Matern.cpp:24-86 ; covparms = (sig2, range, smooth)."""
    d = _f64(distmat)
    out = np.empty_like(d)
    lib().gpv_oracle_MaternFun_flat(d.ravel(), d.size, _f64(covparms), out.ravel())
    return out


def EsqeFun(distmat, covparms):
    """src/Esqe.cpp:17-39 ; covparms = (sig2_1, r1, sig2_2, r2)."""
    d = _f64(distmat)
    out = np.empty_like(d)
    lib().gpv_oracle_EsqeFun_flat(d.ravel(), d.size, _f64(covparms), out.ravel())
    return out


def _colmajor(a, dtype):
    """Return the column-major flat buffer of a 2-D array (what R hands to .Call)."""
    a = np.asarray(a, dtype=dtype)
    return np.ascontiguousarray(a.T).ravel()


def U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType,
                covparms, mode=0):
    """Same nine arguments as the reference's R stub (R/RcppExports.R:22-24).

    locs (N, d) float; revNNarray (N, p) int, 1-based, 0 = missing; revCondOnLatent (N, p) float
    (1/0/NaN).  Returns dict(Lentries=(N, p), Zentries=(2n,), nfail=int).
    mode 0 = LAPACK dpotrf/dtrtrs, 1 = textbook fp64, 2 = __float128 arbiter.
    """
    locs = np.asarray(locs, dtype=np.float64)
    N, d = locs.shape
    revNN = np.asarray(revNNarray)
    p = revNN.shape[1]
    rc = np.asarray(revCondOnLatent, dtype=np.float64)
    L = np.zeros(N * p, dtype=np.float64)
    Z = np.zeros(2 * int(n), dtype=np.float64)
    nfail = C.c_long(0)
    st = lib().gpv_oracle_U_NZentries(int(Ncores), int(n), N, d, p,
                                      _colmajor(locs, np.float64), _colmajor(revNN, np.int32),
                                      _colmajor(rc, np.float64), _f64(nuggets), _f64(nuggets_obsord),
                                      covType.encode(), _f64(covparms), L, Z, int(mode),
                                      C.byref(nfail))
    if st != 0:
        raise ValueError(f"{covType} covariance is not implemented")
    return dict(Lentries=L.reshape(p, N).T.copy(), Zentries=Z, nfail=int(nfail.value))


class RowsProblem:
    """Pre-marshalled inputs for timing the restated reference on a row range (bench.py's CPU
    baseline): all array conversion happens here, `run()` is only the C call."""

    def __init__(self, locs, revNN_rows, revCond_rows, row_begin, nuggets, covType, covparms):
        locs = np.asarray(locs, dtype=np.float64)
        self.N, self.d = locs.shape
        self.nrows, self.p = np.asarray(revNN_rows).shape
        self.row_begin = int(row_begin)
        self.locs = _colmajor(locs, np.float64)
        self.nn = _colmajor(revNN_rows, np.int32)
        rc = np.asarray(revCond_rows)
        if rc.dtype.kind != "f":
            rcf = rc.astype(np.float64)
            rcf[rc < 0] = np.nan
            rc = rcf
        self.rc = _colmajor(rc, np.float64)
        self.nug = _f64(nuggets)
        self.covType = covType.encode()
        self.cov = _f64(covparms)
        self.L = np.zeros(self.nrows * self.p, dtype=np.float64)
        self.Z = np.zeros(0, dtype=np.float64)

    def run(self, threads, mode=0):
        nfail = C.c_long(0)
        st = lib().gpv_oracle_U_NZentries_rows(int(threads), 0, self.N, self.d, self.p, self.row_begin,
                                               self.nrows, self.locs, self.nn, self.rc, self.nug, self.Z,
                                               self.covType, self.cov, self.L, self.Z, int(mode),
                                               C.byref(nfail))
        if st != 0:
            raise ValueError("covariance is not implemented")
        return int(nfail.value)

    def Lentries(self):
        return self.L.reshape(self.p, self.nrows).T


def block_cond_proxy(k, locs, revNNarray, revCondOnLatent, nuggets, covType, covparms):
    locs = np.asarray(locs, dtype=np.float64)
    N, d = locs.shape
    p = np.asarray(revNNarray).shape[1]
    return float(lib().gpv_oracle_block_cond_proxy(
        int(k), N, d, p, _colmajor(locs, np.float64), _colmajor(revNNarray, np.int32),
        _colmajor(revCondOnLatent, np.float64), _f64(nuggets), covType.encode(), _f64(covparms)))


def ic0(ptrs, inds, vals):
    """src/ic0.cpp:43-63 ; ptrs (N+1), inds, vals as R numeric vectors (0-based indices held in
    doubles).  Returns the incomplete-Cholesky values (the input is not modified)."""
    ptrs, inds = _f64(ptrs), _f64(inds)
    out = _f64(vals).copy()
    nerr = lib().gpv_oracle_ic0(ptrs.size - 1, ptrs, inds, out)
    if nerr:
        raise ValueError(f"ic0: {nerr} entries right of the diagonal (the reference prints ERROR)")
    return out


def createUcppM(ptrs, inds, cov_vals):
    """src/ic0.cpp:68-71."""
    return ic0(ptrs, inds, cov_vals)


def createUcpp(ptrs, inds, locsord, covparams, fill_only=False):
    """src/ic0.cpp:77-92 ; locsord (N, d); covparams = (sig2, range, smooth)."""
    ptrs, inds = _f64(ptrs), _f64(inds)
    locs = np.asarray(locsord, dtype=np.float64)
    N, d = locs.shape
    out = np.zeros(inds.size, dtype=np.float64)
    nerr = lib().gpv_oracle_createUcpp(N, d, ptrs, inds, _colmajor(locs, np.float64), _f64(covparams),
                                       int(bool(fill_only)), out)
    if nerr:
        raise ValueError(f"createUcpp: {nerr} entries right of the diagonal")
    return out
