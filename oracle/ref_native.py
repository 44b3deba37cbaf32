"""ctypes front end of oracle/_ref/libgpvecchia_ref.so: the reference's OWN hot-path sources
(/root/reference/src/{U_NZentries,Matern,Esqe,dist,ic0}.cpp, unmodified) compiled against the stand-in
headers of oracle/ref_build/include.  TEST INFRASTRUCTURE ONLY: it pins the restatement
(oracle/ref_c.py) bit for bit and is the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`).

The library is built by `make -C oracle/ref_build` where /root/reference exists (this container); on the
GPU box only the prebuilt file is used (it travels with the snapshot)."""
import ctypes as C
import os
import subprocess

import numpy as np

from .ref_c import _colmajor, _f64, _find_lapack

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgpvecchia_ref.so")
_LIB = None


def available():
    """True when the compiled reference can be loaded (prebuilt, or buildable because /root/reference is here)."""
    return os.path.exists(_SO) or os.path.isdir(os.environ.get("GPV_REFERENCE_SRC", "/root/reference"))


def build():
    ref = os.environ.get("GPV_REFERENCE_SRC", "/root/reference")
    if not os.path.isdir(ref):
        raise FileNotFoundError(f"{ref} not present: oracle/_ref can only be built where the reference sources are")
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref_build"), f"REF={ref}"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.gpv_ref_bind_lapack.argtypes = [C.c_char_p]
    L.gpv_ref_bind_lapack.restype = C.c_int
    L.gpv_ref_has_lapack.restype = C.c_int
    L.gpv_ref_max_threads.restype = C.c_int
    L.gpv_ref_openmp.restype = C.c_int
    L.gpv_ref_force_textbook.argtypes = [C.c_int]
    L.gpv_ref_U_NZentries.argtypes = [C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, dp, ip, dp, dp, dp,
                                      C.c_char_p, dp, C.c_int, dp, dp, C.POINTER(C.c_long)]
    L.gpv_ref_U_NZentries.restype = C.c_int
    L.gpv_ref_U_NZentries_mat.argtypes = [C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, dp, ip, dp, dp, dp,
                                          dp, dp, C.c_int, dp, dp, C.POINTER(C.c_long)]
    L.gpv_ref_U_NZentries_mat.restype = C.c_int
    L.gpv_ref_MaternFun.argtypes = [dp, C.c_long, C.c_long, dp, dp]
    L.gpv_ref_EsqeFun.argtypes = [dp, C.c_long, C.c_long, dp, dp]
    L.gpv_ref_ic0.argtypes = [C.c_long, dp, dp, C.c_long, dp]
    L.gpv_ref_ic0.restype = C.c_long
    L.gpv_ref_createUcppM.argtypes = [C.c_long, dp, dp, C.c_long, dp]
    L.gpv_ref_createUcppM.restype = C.c_long
    L.gpv_ref_createUcpp.argtypes = [C.c_long, C.c_int, dp, dp, C.c_long, dp, dp, dp]
    L.gpv_ref_createUcpp.restype = C.c_long
    lp = _find_lapack()
    if lp is not None:
        L.gpv_ref_bind_lapack(lp.encode())
    _LIB = L
    return L


def has_lapack():
    return bool(lib().gpv_ref_has_lapack())


def max_threads():
    return int(lib().gpv_ref_max_threads())


def force_textbook(on):
    """chol/solve by the published unblocked algorithms instead of OpenBLAS (timed beside it in bench.py)."""
    lib().gpv_ref_force_textbook(int(bool(on)))


def MaternFun(distmat, covparms):
    """The reference's MaternFun (src/Matern.cpp:24-86) on a matrix of distances."""
    d = np.atleast_2d(np.asarray(distmat, dtype=np.float64))
    out = np.empty(d.size, dtype=np.float64)
    lib().gpv_ref_MaternFun(_colmajor(d, np.float64), d.shape[0], d.shape[1], _f64(covparms), out)
    return out.reshape(d.shape[1], d.shape[0]).T.copy()


def EsqeFun(distmat, covparms):
    """The reference's EsqeFun (src/Esqe.cpp:17-39)."""
    d = np.atleast_2d(np.asarray(distmat, dtype=np.float64))
    out = np.empty(d.size, dtype=np.float64)
    lib().gpv_ref_EsqeFun(_colmajor(d, np.float64), d.shape[0], d.shape[1], _f64(covparms), out)
    return out.reshape(d.shape[1], d.shape[0]).T.copy()


class Problem:
    """Inputs of one U_NZentries call marshalled once (column-major, as R holds them); `run()` is only the
    call into the compiled reference.  Used by bench.py to time it."""

    def __init__(self, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType, covparms):
        locs = np.asarray(locs, dtype=np.float64)
        self.N, self.d = locs.shape
        self.p = np.asarray(revNNarray).shape[1]
        self.n = int(n)
        self.locs = _colmajor(locs, np.float64)
        self.nn = _colmajor(revNNarray, np.int32)
        rc = np.asarray(revCondOnLatent)
        if rc.dtype.kind != "f":                 # R logical (NA = INT_MIN) -> double (NA -> NaN), as Rcpp coerces it
            rcf = rc.astype(np.float64)
            rcf[rc < 0] = np.nan
            rc = rcf
        self.rc = _colmajor(rc, np.float64)
        self.nug = _f64(nuggets)
        self.nug_obs = _f64(nuggets_obsord)
        self.covType = covType.encode()
        self.cov = _f64(covparms)
        self.L = np.zeros(self.N * self.p, dtype=np.float64)
        self.Z = np.zeros(2 * self.n, dtype=np.float64)
        self.nmsg = 0

    def run(self, Ncores):
        nmsg = C.c_long(0)
        st = lib().gpv_ref_U_NZentries(int(Ncores), self.n, self.N, self.d, self.p, self.locs, self.nn, self.rc,
                                       self.nug, self.nug_obs, self.covType, self.cov, self.cov.size, self.L,
                                       self.Z, C.byref(nmsg))
        self.nmsg = int(nmsg.value)
        return st

    def Lentries(self):
        return self.L.reshape(self.p, self.N).T


def U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType, covparms):
    """The reference's U_NZentries with its nine arguments (R/RcppExports.R:22-24).  Returns
    dict(Lentries (N, p), Zentries (2n,), nfail = messages written to Rcerr)."""
    pr = Problem(n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType, covparms)
    st = pr.run(Ncores)
    if st != 0:
        raise ValueError(f"the reference threw (covType {covType!r})")
    return dict(Lentries=pr.Lentries().copy(), Zentries=pr.Z, nfail=pr.nmsg)


def U_NZentries_mat(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covVals, covparms):
    """The reference's U_NZentries_mat (src/U_NZentries.cpp:126-197)."""
    locs = np.asarray(locs, dtype=np.float64)
    N, d = locs.shape
    p = np.asarray(revNNarray).shape[1]
    L = np.zeros(N * p, dtype=np.float64)
    Z = np.zeros(2 * int(n), dtype=np.float64)
    nmsg = C.c_long(0)
    cp = _f64(covparms)
    st = lib().gpv_ref_U_NZentries_mat(int(Ncores), int(n), N, d, p, _colmajor(locs, np.float64),
                                       _colmajor(revNNarray, np.int32),
                                       _colmajor(np.asarray(revCondOnLatent, dtype=np.float64), np.float64),
                                       _f64(nuggets), _f64(nuggets_obsord), _colmajor(covVals, np.float64), cp,
                                       cp.size, L, Z, C.byref(nmsg))
    if st != 0:
        raise ValueError("the reference threw")
    return dict(Lentries=L.reshape(p, N).T.copy(), Zentries=Z, nfail=int(nmsg.value))


def ic0(ptrs, inds, vals):
    """The reference's ic0 (src/ic0.cpp:43-63); returns the new values (the input is copied first)."""
    ptrs, inds = _f64(ptrs).copy(), _f64(inds).copy()
    out = _f64(vals).copy()
    nerr = lib().gpv_ref_ic0(ptrs.size - 1, ptrs, inds, out.size, out)
    if nerr:
        raise ValueError(f"ic0: the reference printed ERROR {nerr} times")
    return out


def createUcppM(ptrs, inds, cov_vals):
    ptrs, inds = _f64(ptrs).copy(), _f64(inds).copy()
    out = _f64(cov_vals).copy()
    lib().gpv_ref_createUcppM(ptrs.size - 1, ptrs, inds, out.size, out)
    return out


def createUcpp(ptrs, inds, locsord, covparams):
    """The reference's createUcpp (src/ic0.cpp:77-92)."""
    ptrs, inds = _f64(ptrs).copy(), _f64(inds).copy()
    locs = np.asarray(locsord, dtype=np.float64)
    N, d = locs.shape
    out = np.zeros(inds.size, dtype=np.float64)
    lib().gpv_ref_createUcpp(N, d, ptrs, inds, inds.size, _colmajor(locs, np.float64), _f64(covparams), out)
    return out
