"""Host-side mirror of the reference's interface for the createU / U_NZentries path.

Same names, argument meaning and error behaviour as the reference (GPvecchia 0.1.8):
  U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType,
              covparms)                      R/RcppExports.R:22-24, src/U_NZentries.cpp:25-118
  MaternFun(distmat, covparms)               R/RcppExports.R:15-17, src/Matern.cpp:24
  EsqeFun(distmat, covparms)                 R/RcppExports.R:4-6,   src/Esqe.cpp:17
  U_sparsity(locs, NNarray, obs, Cond)       R/U_sparsity.R:5-81   (vectorised, same arrays)
  createU(vecchia.approx, covparms, nuggets, covmodel)            R/createU.R:65-201
  vecchia_likelihood(z, vecchia.approx, covparms, nuggets, covmodel)  R/vecchia_likelihood.R:14-27
The reference's host language is R, which this image does not have; the R-side `.Call` shim a
maintainer would add is in r_shim/ and INTEGRATION.md.  This Python mirror drives the SAME C ABI
(include/gpvecchia_b200.h) through ctypes, so tests read like the reference's own.

All numbers come from the CUDA library; NumPy/SciPy are used only for what stays on the host in
the reference as well (argument preparation, Matrix::sparseMatrix assembly, the denominator's
sparse algebra).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _lib
from ._lib import lib, check

NA_INT = 0          # NA in (rev)NNarray, as createU.R:146-147 passes it to C++
R_NA_LOGICAL = np.iinfo(np.int32).min


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _colmajor(a, dtype):
    """Flat column-major copy of a 2-D array: the buffer R hands to .Call."""
    return np.ascontiguousarray(np.asarray(a, dtype=dtype).T).ravel()


def _nn_to_i32(revNNarray):
    a = np.asarray(revNNarray)
    if a.dtype.kind == "f":
        a = np.where(np.isnan(a), 0, a)
    return a.astype(np.int32)


def _cond_to_rlogical(revCond):
    """bool / int8 (-1 = NA) / float (NaN = NA) -> R logical storage (int32, NA = INT_MIN)."""
    a = np.asarray(revCond)
    out = np.empty(a.shape, dtype=np.int32)
    if a.dtype.kind == "f":
        nan = np.isnan(a)
        out[...] = np.where(nan, 0, a).astype(np.int32)
        out[nan] = R_NA_LOGICAL
    elif a.dtype.kind == "b":
        out[...] = a.astype(np.int32)
    else:
        out[...] = a.astype(np.int32)
        out[a < 0] = R_NA_LOGICAL
    return out


class UHandle:
    """Device-resident, parameter-free part of one vecchia.approx (gpv_create).

    createU is called many times per vecchia_specify with only covparms/nuggets changing
    (vecchia_estimate's optimiser, the VL Newton loop), so locsord / revNNarray / revCond / obs
    are uploaded once and stay in HBM.
    """

    def __init__(self, locsord, revNNarray, revCond, obs=None, row_begin=0, row_end=None, device=0):
        """revNNarray / revCond: either all N rows (as R holds them) or only the rows
        [row_begin,row_end) of this shard (gpv_create_shard)."""
        locs = np.asarray(locsord, dtype=np.float64)
        if locs.ndim != 2:
            raise ValueError("Locations must be in matrix form")   # vecchia_specify.R:32-35
        self.N, self.d = locs.shape
        nn = _nn_to_i32(revNNarray)
        self.p = nn.shape[1]
        self.row_begin = int(row_begin)
        self.row_end = self.N if row_end is None else int(row_end)
        self.nrows = self.row_end - self.row_begin
        self._shard_arrays = nn.shape[0] != self.N
        if self._shard_arrays and nn.shape[0] != self.nrows:
            raise ValueError("revNNarray must have one row per location, or one per row of the shard")
        if np.asarray(revCond).shape != nn.shape:
            raise ValueError("revCond must have the shape of revNNarray")
        self.device = int(device)
        self.n_obs = 0
        obs_i32 = None
        if obs is not None:
            obs_i32 = np.ascontiguousarray(np.asarray(obs).astype(bool).astype(np.int32))
            self.n_obs = int(obs_i32.sum())
        h = C.c_void_p()
        create = lib.gpv_create_shard if self._shard_arrays else lib.gpv_create
        check(create(C.byref(h), self.N, self.p, self.d, _ptr(_colmajor(locs, np.float64)),
                             _ptr(_colmajor(nn, np.int32)),
                             _ptr(_colmajor(_cond_to_rlogical(revCond), np.int32)),
                             _lib.GPV_COND_RLOGICAL_I32, _ptr(obs_i32), self.row_begin, self.row_end,
                             self.device))
        self._h = h
        self.packed_len = int(lib.gpv_packed_len(h))
        self.nuggets_read = int(lib.gpv_nuggets_read(h))    # leading entries of `nuggets` a call uploads

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.gpv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- calls ---------------------------------------------------------------------------------
    def set_revcond(self, revCond):
        check(lib.gpv_set_revcond(self._h, _ptr(_colmajor(_cond_to_rlogical(revCond), np.int32)),
                                  _lib.GPV_COND_RLOGICAL_I32))

    def U_NZentries(self, covType, covparms, nuggets, nuggets_obsord):
        """list(Lentries = nrows x p, Zentries = 2n) like U_NZentries.cpp:117, plus fail info."""
        cov = _f64(covparms)
        nug = _f64(nuggets)
        if nug.size != self.N:
            raise ValueError("nuggets must have one entry per location")
        tau = _f64(nuggets_obsord)
        n = tau.size
        L = np.empty(self.nrows * self.p, dtype=np.float64)
        Z = np.empty(2 * n, dtype=np.float64)
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_u_nzentries(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug), _ptr(tau),
                                  n, _ptr(L), _ptr(Z), C.byref(nfail), C.byref(first)))
        return dict(Lentries=L.reshape(self.p, self.nrows).T, Zentries=Z, nfail=int(nfail.value),
                    first_fail=int(first.value))

    def U_NZentries_mat(self, covVals, nuggets_obsord):
        """U_NZentries_mat (src/U_NZentries.cpp:126-197): `covmodel` given as the N x N covariance matrix."""
        cv = np.asfortranarray(np.asarray(covVals, dtype=np.float64))
        if cv.shape != (self.N, self.N):
            raise ValueError("covVals must be Nlocs x Nlocs")
        tau = _f64(nuggets_obsord)
        n = tau.size
        L = np.empty(self.nrows * self.p, dtype=np.float64)
        Z = np.empty(2 * n, dtype=np.float64)
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_u_nzentries_mat(self._h, cv.ctypes.data_as(C.c_void_p), _ptr(tau), n, _ptr(L), _ptr(Z),
                                      C.byref(nfail), C.byref(first)))
        return dict(Lentries=L.reshape(self.p, self.nrows).T, Zentries=Z, nfail=int(nfail.value),
                    first_fail=int(first.value))

    def values_packed_mat(self, covVals, nuggets_obsord, zentries_tail=True):
        cv = np.asfortranarray(np.asarray(covVals, dtype=np.float64))
        if cv.shape != (self.N, self.N):
            raise ValueError("covVals must be Nlocs x Nlocs")
        tau = _f64(nuggets_obsord)
        n = tau.size
        out = np.empty(self.packed_len + (2 * n if zentries_tail else 0), dtype=np.float64)
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_u_values_packed_mat(self._h, cv.ctypes.data_as(C.c_void_p), _ptr(tau), n,
                                          1 if zentries_tail else 0, _ptr(out), C.byref(nfail), C.byref(first)))
        return out, int(nfail.value), int(first.value)

    def values_packed(self, covType, covparms, nuggets, nuggets_obsord, zentries_tail=True, out=None):
        """allLentries of createU.R:158-160 for this shard, straight from the device.  nuggets = nuggets_obsord
        = None: the vectors set_scalar_nugget() left on the device."""
        cov = _f64(covparms)
        nug = None if nuggets is None else _f64(nuggets)
        tau = None if nuggets_obsord is None else _f64(nuggets_obsord)
        n = self.n_obs if tau is None else tau.size
        total = self.packed_len + (2 * n if zentries_tail else 0)
        if out is None:
            out = np.empty(total, dtype=np.float64)
        elif out.size < total or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of gpv_packed_len (+2n)")
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_u_values_packed(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug),
                                      _ptr(tau), n, 1 if zentries_tail else 0, _ptr(out),
                                      C.byref(nfail), C.byref(first)))
        return out, int(nfail.value), int(first.value)

    def csc_dims(self):
        """(ncols, nnz, size): U columns covered by this shard, their nonzeros, N + n."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(lib.gpv_csc_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def u_sparsity(self):
        """colindices, rowpointers of R/U_sparsity.R:36-73 (1-based int32) for this shard's rows."""
        _, nnz, _ = self.csc_dims()
        ci, rp = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.int32)
        check(lib.gpv_u_sparsity(self._h, _ptr(ci), _ptr(rp)))
        return ci, rp

    def csc_pattern(self):
        """dgCMatrix@p, @i (0-based int32) of the U columns of this shard's rows."""
        ncols, nnz, _ = self.csc_dims()
        colptr, rowidx = np.empty(ncols + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32)
        check(lib.gpv_u_csc_pattern(self._h, _ptr(colptr), _ptr(rowidx)))
        return colptr, rowidx

    def values_csc(self, covType, covparms, nuggets, nuggets_obsord, out=None):
        """dgCMatrix@x of the U columns of this shard's rows (createU.R:152-161 in one call).  nuggets =
        nuggets_obsord = None: the vectors set_scalar_nugget() left on the device."""
        cov = _f64(covparms)
        nug = None if nuggets is None else _f64(nuggets)
        tau = None if nuggets_obsord is None else _f64(nuggets_obsord)
        n = self.n_obs if tau is None else tau.size
        _, nnz, _ = self.csc_dims()
        if out is None:
            out = np.empty(nnz, dtype=np.float64)
        elif out.size < nnz or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of nnz doubles")
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_u_values_csc(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug), _ptr(tau), n,
                                   _ptr(out), C.byref(nfail), C.byref(first)))
        return out, int(nfail.value), int(first.value)

    def set_scalar_nugget(self, nugget):
        """nuggets.all.ord / nuggets.ord of a scalar nugget, built on the device (createU.R:70-78); the
        likelihood and U-values calls then take nuggets=None."""
        check(lib.gpv_set_scalar_nugget(self._h, float(nugget)))

    def loglik_numerator(self, covType, covparms, nuggets, nuggets_obsord, zord, skip_rows=0,
                         include_obs_terms=-1):
        """(quadform.num, logdet.num, nfail) of vecchia_likelihood.R:74-76, fused on the GPU.  None for the
        nuggets / for zord reuses what is resident on the handle (estimation loop)."""
        cov = _f64(covparms)
        nug = None if nuggets is None else _f64(nuggets)
        tau = None if nuggets_obsord is None else _f64(nuggets_obsord)
        z = None if zord is None else _f64(zord)
        out = np.zeros(3, dtype=np.float64)
        check(lib.gpv_loglik_numerator(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug),
                                       _ptr(tau), _ptr(z), self.n_obs, int(skip_rows),
                                       int(include_obs_terms), _ptr(out)))
        return float(out[0]), float(out[1]), int(out[2])

    def loglik_z(self, covType, covparms, nuggets, nuggets_obsord, zord, include_obs_terms=-1):
        """Whole log-likelihood for pure `z` conditioning (gpv_loglik_z): dict(loglik, quadform_num,
        logdet_num, quadform_denom, logdet_denom, nfail).  None arguments reuse resident data."""
        cov = _f64(covparms)
        nug = None if nuggets is None else _f64(nuggets)
        tau = None if nuggets_obsord is None else _f64(nuggets_obsord)
        z = None if zord is None else _f64(zord)
        out = np.zeros(6, dtype=np.float64)
        check(lib.gpv_loglik_z(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug), _ptr(tau), _ptr(z),
                               self.n_obs, int(include_obs_terms), _ptr(out)))
        return dict(loglik=float(out[0]), quadform_num=float(out[1]), logdet_num=float(out[2]),
                    quadform_denom=float(out[3]), logdet_denom=float(out[4]), nfail=int(out[5]))

    # ---- multi-process runs: slices exchanged over NVLink (NCCL), partial sums all-reduced (gpv_dist.inc) ----
    @staticmethod
    def dist_unique_id():
        """128 bytes from rank 0 (ncclGetUniqueId) that the caller carries to every rank."""
        buf = np.zeros(128, dtype=np.uint8)
        check(lib.gpv_dist_unique_id(_ptr(buf)))
        return buf

    def dist_init(self, unique_id, rank, world):
        uid = np.ascontiguousarray(np.asarray(unique_id, dtype=np.uint8))
        if uid.size != 128:
            raise ValueError("unique_id must be the 128 bytes of dist_unique_id()")
        check(lib.gpv_dist_init(self._h, _ptr(uid), int(rank), int(world)))
        self._dist = (int(rank), int(world))

    def loglik_z_dist(self, covType, covparms, nuggets_slice, tau_slice, z_slice, loc_cuts, obs_cuts):
        """Whole log-likelihood over all ranks from each rank's SLICE of nuggets.all.ord / nuggets.ord / zord
        (gpv_loglik_z_dist); every rank returns the complete value.  None slices reuse the previous call's data."""
        cov = _f64(covparms)
        ns = None if nuggets_slice is None else _f64(nuggets_slice)
        ts = None if tau_slice is None else _f64(tau_slice)
        zs = None if z_slice is None else _f64(z_slice)
        lc = None if loc_cuts is None else np.ascontiguousarray(np.asarray(loc_cuts, dtype=np.int64))
        oc = None if obs_cuts is None else np.ascontiguousarray(np.asarray(obs_cuts, dtype=np.int64))
        out = np.zeros(6, dtype=np.float64)
        check(lib.gpv_loglik_z_dist(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(ns), _ptr(ts), _ptr(zs),
                                    _ptr(lc), _ptr(oc), _ptr(out)))
        return dict(loglik=float(out[0]), quadform_num=float(out[1]), logdet_num=float(out[2]),
                    quadform_denom=float(out[3]), logdet_denom=float(out[4]), nfail=int(out[5]))

    def u_dev(self, covType, covparms, d_nuggets, d_out=None, packed=False, d_zord=None, skip_rows=0,
              d_loglik=None, stream=None):
        """Device-pointer variant (ints are raw device addresses, e.g. torch_tensor.data_ptr())."""
        cov = _f64(covparms)
        check(lib.gpv_u_dev(self._h, covType.encode(), _ptr(cov), cov.size, C.c_void_p(d_nuggets),
                            C.c_void_p(d_out) if d_out else None, 1 if packed else 0,
                            C.c_void_p(d_zord) if d_zord else None, int(skip_rows),
                            C.c_void_p(d_loglik) if d_loglik else None,
                            C.c_void_p(stream) if stream else None))

    def last_kernel_ms(self):
        ms = C.c_float(0)
        check(lib.gpv_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_time_stats(self, reset=True):
        """(launch count, total ms) of the set kernel since the last reset, from per-launch CUDA
        events recorded on the launching stream inside the library."""
        cnt, tot = C.c_int64(0), C.c_double(0.0)
        check(lib.gpv_kernel_time_stats(self._h, 1 if reset else 0, C.byref(cnt), C.byref(tot)))
        return int(cnt.value), float(tot.value)

    def last_kernel_name(self):
        return lib.gpv_last_kernel_name(self._h).decode()


class MultiHandle:
    """One process, several GPUs (gpv_multi_*): what an R session would hold.  Rows are split into
    contiguous ranges balancing sum n0^3; every device writes its slice of the packed vector."""

    def __init__(self, locsord, revNNarray, revCond, obs=None, devices=(0,)):
        locs = np.asarray(locsord, dtype=np.float64)
        self.N, self.d = locs.shape
        nn = _nn_to_i32(revNNarray)
        self.p = nn.shape[1]
        obs_i32 = None
        self.n_obs = 0
        if obs is not None:
            obs_i32 = np.ascontiguousarray(np.asarray(obs).astype(bool).astype(np.int32))
            self.n_obs = int(obs_i32.sum())
        dev = np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
        h = C.c_void_p()
        check(lib.gpv_multi_create(C.byref(h), self.N, self.p, self.d, _ptr(_colmajor(locs, np.float64)),
                                   _ptr(_colmajor(nn, np.int32)),
                                   _ptr(_colmajor(_cond_to_rlogical(revCond), np.int32)),
                                   _lib.GPV_COND_RLOGICAL_I32, _ptr(obs_i32), _ptr(dev), dev.size))
        self._h = h
        self.packed_len = int(lib.gpv_multi_packed_len(h))
        cuts = np.zeros(dev.size + 1, dtype=np.int64)
        lib.gpv_multi_row_cuts(h, _ptr(cuts))
        self.row_cuts = cuts

    def close(self):
        if getattr(self, "_h", None):
            lib.gpv_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def values_packed(self, covType, covparms, nuggets, nuggets_obsord, zentries_tail=True, out=None):
        cov, nug, tau = _f64(covparms), _f64(nuggets), _f64(nuggets_obsord)
        n = tau.size
        total = self.packed_len + (2 * n if zentries_tail else 0)
        if out is None:
            out = np.empty(total, dtype=np.float64)
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_multi_u_values_packed(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug),
                                            _ptr(tau), n, 1 if zentries_tail else 0, _ptr(out),
                                            C.byref(nfail), C.byref(first)))
        return out, int(nfail.value), int(first.value)

    def csc_dims(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(lib.gpv_multi_csc_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def csc_pattern(self):
        ncols, nnz, _ = self.csc_dims()
        colptr, rowidx = np.empty(ncols + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32)
        check(lib.gpv_multi_u_csc_pattern(self._h, _ptr(colptr), _ptr(rowidx)))
        return colptr, rowidx

    def values_csc(self, covType, covparms, nuggets, nuggets_obsord, out=None):
        cov, nug, tau = _f64(covparms), _f64(nuggets), _f64(nuggets_obsord)
        _, nnz, _ = self.csc_dims()
        if out is None:
            out = np.empty(nnz, dtype=np.float64)
        nfail, first = C.c_int64(0), C.c_int64(-1)
        check(lib.gpv_multi_u_values_csc(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug), _ptr(tau),
                                         tau.size, _ptr(out), C.byref(nfail), C.byref(first)))
        return out, int(nfail.value), int(first.value)

    def loglik_numerator(self, covType, covparms, nuggets, nuggets_obsord, zord, skip_rows=0):
        cov, nug, tau, z = _f64(covparms), _f64(nuggets), _f64(nuggets_obsord), _f64(zord)
        out = np.zeros(3, dtype=np.float64)
        check(lib.gpv_multi_loglik_numerator(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug),
                                             _ptr(tau), _ptr(z), tau.size, int(skip_rows), _ptr(out)))
        return float(out[0]), float(out[1]), int(out[2])

    def loglik_z(self, covType, covparms, nuggets, nuggets_obsord, zord):
        cov, nug, tau, z = _f64(covparms), _f64(nuggets), _f64(nuggets_obsord), _f64(zord)
        out = np.zeros(6, dtype=np.float64)
        check(lib.gpv_multi_loglik_z(self._h, covType.encode(), _ptr(cov), cov.size, _ptr(nug), _ptr(tau),
                                     _ptr(z), tau.size, _ptr(out)))
        return dict(loglik=float(out[0]), quadform_num=float(out[1]), logdet_num=float(out[2]),
                    quadform_denom=float(out[3]), logdet_denom=float(out[4]), nfail=int(out[5]))


# --------------------------------------------------------------------------------------------------
# the reference's stateless entry points
# --------------------------------------------------------------------------------------------------
def U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord, covType,
                covparms, device=0):
    """Drop-in for the reference's U_NZentries: nine arguments, list(Lentries, Zentries)."""
    locs = np.asarray(locs, dtype=np.float64)
    N, d = locs.shape
    nn = _nn_to_i32(revNNarray)
    p = nn.shape[1]
    cov = _f64(covparms)
    nug = _f64(nuggets)
    tau = _f64(nuggets_obsord)
    if tau.size != int(n):
        raise ValueError("length(nuggets_obsord) must equal n")
    L = np.empty(N * p, dtype=np.float64)
    Z = np.empty(2 * int(n), dtype=np.float64)
    nfail, first = C.c_int64(0), C.c_int64(-1)
    check(lib.gpv_U_NZentries(int(Ncores), int(n), N, p, d, _ptr(_colmajor(locs, np.float64)),
                              _ptr(_colmajor(nn, np.int32)),
                              _ptr(_colmajor(_cond_to_rlogical(revCondOnLatent), np.int32)),
                              _lib.GPV_COND_RLOGICAL_I32, _ptr(nug), _ptr(tau), covType.encode(),
                              _ptr(cov), cov.size, _ptr(L), _ptr(Z), C.byref(nfail), C.byref(first),
                              int(device)))
    return dict(Lentries=L.reshape(p, N).T, Zentries=Z, nfail=int(nfail.value),
                first_fail=int(first.value))


def MaternFun(distmat, covparms, device=0):
    d = _f64(distmat)
    out = np.empty_like(d)
    cov = _f64(covparms)
    if cov.size != 3:
        raise ValueError("covparms = c(sig2, range, smooth)")
    check(lib.gpv_MaternFun(_ptr(d), d.size, _ptr(cov), _ptr(out), int(device)))
    return out


def EsqeFun(distmat, covparms, device=0):
    d = _f64(distmat)
    out = np.empty_like(d)
    cov = _f64(covparms)
    if cov.size != 4:
        raise ValueError("covparms = c(sig2_1, r1, sig2_2, r2)")
    check(lib.gpv_EsqeFun(_ptr(d), d.size, _ptr(cov), _ptr(out), int(device)))
    return out


# --------------------------------------------------------------------------------------------------
# ic0 / createUcppM / createUcpp (R/RcppExports.R:53-63): the incomplete-Cholesky branch of createU
# --------------------------------------------------------------------------------------------------
def ic0(ptrs, inds, vals):
    """Incomplete Cholesky on a fixed pattern (src/ic0.cpp:43-63): ptrs (N+1), inds and vals as R numeric
    vectors, 0-based.  Returns the new values; a float64 array passed as `vals` is overwritten in place,
    as the reference overwrites its argument."""
    ptrs, inds = _f64(ptrs), _f64(inds)
    out = vals if (isinstance(vals, np.ndarray) and vals.dtype == np.float64 and vals.flags.c_contiguous) else _f64(vals).copy()
    if out.size != inds.size:
        raise ValueError("vals and inds differ in length")
    check(lib.gpv_ic0(ptrs.size - 1, _ptr(ptrs), _ptr(inds), inds.size, _ptr(out)))
    return out


def createUcppM(ptrs, inds, cov_vals):
    """src/ic0.cpp:68-71: ic0 on covariances the caller evaluated (matrix or function covmodel)."""
    ptrs, inds = _f64(ptrs), _f64(inds)
    out = cov_vals if (isinstance(cov_vals, np.ndarray) and cov_vals.dtype == np.float64 and cov_vals.flags.c_contiguous) else _f64(cov_vals).copy()
    if out.size != inds.size:
        raise ValueError("cov_vals and inds differ in length")
    check(lib.gpv_createUcppM(ptrs.size - 1, _ptr(ptrs), _ptr(inds), inds.size, _ptr(out)))
    return out


def createUcpp(ptrs, inds, locsord, covparams, device=0):
    """src/ic0.cpp:77-92: Matern covariance of every stored (row, column) pair on the device, then ic0."""
    ptrs, inds = _f64(ptrs), _f64(inds)
    locs = np.asarray(locsord, dtype=np.float64)
    if locs.ndim != 2 or locs.shape[0] != ptrs.size - 1:
        raise ValueError("locsord must have one row per pattern row")
    cov = _f64(covparams)
    if cov.size != 3:
        raise ValueError("covparams = c(sig2, range, smooth)")
    lc = _colmajor(locs, np.float64)
    out = np.zeros(inds.size, dtype=np.float64)
    check(lib.gpv_createUcpp(locs.shape[0], locs.shape[1], _ptr(ptrs), _ptr(inds), inds.size, _ptr(lc),
                             _ptr(cov), _ptr(out), int(device)))
    return out


# --------------------------------------------------------------------------------------------------
# U_sparsity, vectorised (R/U_sparsity.R:5-81 is an interpreter loop over all locations)
# --------------------------------------------------------------------------------------------------
def U_sparsity(locs, NNarray, obs, Cond):
    """Same outputs as the reference's U_sparsity (1-based ids, 0 for NA in NNarray, Cond as
    int8 1/0/-1).  Bit-identical arrays for NN layouts; tests compare with the loop restatement."""
    NNarray = np.asarray(NNarray, dtype=np.int64)
    Cond = np.asarray(Cond, dtype=np.int8)
    obs = np.asarray(obs, dtype=bool)
    nnp = locs.shape[0]
    n = int(obs.sum())
    size = nnp + n
    before = np.cumsum(obs) - obs                       # observed locations strictly before k
    latent_map = np.arange(1, nnp + 1, dtype=np.int64) + before
    observed_map = np.where(obs, latent_map + 1, 0)
    revNNarray = NNarray[:, ::-1].copy()
    revCond = Cond[:, ::-1].copy()
    keep = revNNarray != NA_INT
    rows = np.broadcast_to(latent_map[:, None], revNNarray.shape)[keep]
    ids = revNNarray[keep]
    cnd = revCond[keep] == 1
    cols = np.where(cnd, latent_map[ids - 1], observed_map[ids - 1])
    ok = np.nonzero(obs)[0]
    Zrow = np.repeat(observed_map[ok], 2)
    Zcol = np.stack([latent_map[ok], observed_map[ok]], axis=1).ravel()
    return dict(revNNarray=revNNarray, revCond=revCond, n_cores=1, size=size,
                rowpointers=np.concatenate([rows, Zrow]), colindices=np.concatenate([cols, Zcol]),
                y_ind=latent_map, observed_map=observed_map)


# --------------------------------------------------------------------------------------------------
# createU (R/createU.R:65-201, non-MRA branch) on top of the device handle
# --------------------------------------------------------------------------------------------------
def _handle_for(va, device=0):
    h = va.get("_gpv_handle")
    if h is None or h._h is None or h.device != device:
        prep = va["U_prep"]
        h = UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"], device=device)
        va["_gpv_handle"] = h     # like an R external pointer: re-created lazily if absent
    return h


def _prepare_nuggets(va, nuggets):
    """createU.R:67-80."""
    prep = va["U_prep"]
    obs = np.asarray(va["obs"], dtype=bool)
    n = int(obs.sum())
    size = prep["size"]
    latent = np.isin(np.arange(1, size + 1), prep["y_ind"])
    ord_ = np.asarray(va["ord"])
    nuggets = np.atleast_1d(np.asarray(nuggets, dtype=np.float64))
    if nuggets.size == 1:
        nuggets = np.repeat(nuggets, n)
    nuggets_all = np.concatenate([nuggets, np.zeros(int(latent.sum()) - n)])
    ord_all = np.concatenate([ord_[:n], ord_ + n]) if va["cond_yz"] == "zy" else ord_
    nuggets_all_ord = nuggets_all[ord_all - 1]
    nuggets_ord = nuggets_all[np.asarray(va["ord_z"]) - 1]
    return n, size, latent, ord_, obs, nuggets, nuggets_all_ord, nuggets_ord


def createU(vecchia_approx, covparms, nuggets, covmodel="matern", device=0, assemble="csc"):
    """createU (R/createU.R:65-201, non-MRA branch).  assemble = "csc": the device delivers the slots of
    the compressed-column matrix (gpv_u_values_csc / gpv_u_csc_pattern); "triplet": the reference's route,
    packed values + sparseMatrix(i, j, x) (:158-161).  Same matrix either way."""
    va = vecchia_approx
    prep = va["U_prep"]
    if va.get("conditioning", "NN") == "mra":
        raise NotImplementedError("the MRA / ic0 branch (createU.R:89-139) stays in the reference")
    covmat = None
    if isinstance(covmodel, np.ndarray) and covmodel.ndim == 2:        # createU.R:149-151: U_NZentries_mat
        covmat, assemble = covmodel, "triplet"
    elif not isinstance(covmodel, str):
        raise TypeError("argument 'covmodel' type not supported")      # createU.R:155
    scalar_nugget = np.ndim(nuggets) == 0 or np.size(nuggets) == 1
    n, size, latent, ord_, obs, nuggets, nuggets_all_ord, nuggets_ord = _prepare_nuggets(va, nuggets)
    zero_nuggets = bool(np.any(nuggets == 0))
    h = _handle_for(va, device)
    if scalar_nugget and not zero_nuggets and covmat is None and va["cond_yz"] != "zy" and n > 0:
        # one scalar crosses the boundary instead of two vectors: the device builds nuggets.all.ord and
        # nuggets.ord itself (createU.R:70-78: the nugget at observed locations, 0 elsewhere)
        h.set_scalar_nugget(float(nuggets[0]))
        nuggets_all_ord = nuggets_ord = None
    restore = False
    if zero_nuggets:                                                   # createU.R:83-86
        revCond = prep["revCond"].copy()
        zero_ids = np.nonzero(nuggets_ord == 0)[0] + 1
        revCond[np.isin(prep["revNNarray"], zero_ids) & (prep["revNNarray"] != NA_INT)] = 1
        h.set_revcond(revCond)
        restore = True
    U = None
    try:
        if assemble == "csc":
            try:
                colptr, rowidx = h.csc_pattern()
                x, nfail, first_fail = h.values_csc(covmodel, covparms, nuggets_all_ord, nuggets_ord)
                U = sp.csc_matrix((x, rowidx, colptr), shape=(size, size))
            except _lib.GpvError as e:
                if e.status != _lib.GPV_ERR_UNSUPPORTED:
                    raise
        if U is None and covmat is not None:
            allLentries, nfail, first_fail = h.values_packed_mat(covmat, nuggets_ord)
        elif U is None:
            # the device writes allLentries = c(c(t(Lentries))[not.na], Zentries) directly (:158-160)
            allLentries, nfail, first_fail = h.values_packed(covmodel, covparms, nuggets_all_ord, nuggets_ord)
    finally:
        if restore:
            h.set_revcond(prep["revCond"])
    if nfail:
        import warnings
        warnings.warn(f"Cholesky decomposition failed for {nfail} conditioning set(s) "
                      f"(first at row {first_fail + 1}); those rows of U are zero")
    if U is None:
        U = sp.coo_matrix((allLentries, (prep["colindices"] - 1, prep["rowpointers"] - 1)),
                          shape=(size, size)).tocsc()                  # :161-162
        U.sum_duplicates()
    if va["cond_yz"] == "zy":                                          # :166-171
        keep = np.ones(size, dtype=bool)
        keep[2 * np.arange(n)] = False
        U = U[keep][:, keep]
        latent = latent[keep]
        keep_obs = np.ones(obs.size, dtype=bool)
        keep_obs[n:2 * n] = False
        obs = obs[keep_obs]
    zero_nugg = {}
    if zero_nuggets:                                                   # :174-193
        if va["cond_yz"] == "zy":
            raise NotImplementedError("zy + zero nuggets relies on R recycling semantics")
        diagU = U.diagonal()
        inds_U = np.nonzero(np.isinf(diagU) & (diagU > 0))[0] + 1
        Uc = U.tocsc()
        cond_on = np.array([Uc.indices[Uc.indptr[j - 1]:Uc.indptr[j]][
            Uc.data[Uc.indptr[j - 1]:Uc.indptr[j]] != 0].min() + 1 for j in inds_U], dtype=np.int64)
        keep = np.ones(U.shape[0], dtype=bool)
        keep[inds_U - 1] = False
        U = U[keep][:, keep]
        all_idx = np.arange(1, size + 1)
        inds_z = np.nonzero(np.isin(all_idx[~latent], inds_U))[0] + 1
        inds_locs = np.nonzero(np.isin(all_idx[latent], cond_on))[0] + 1
        zero_nugg = dict(inds_U=inds_U, inds_z=inds_z, inds_locs=inds_locs)
        latent = latent.copy()
        latent[cond_on - 1] = False
        latent = latent[keep]
        sel = np.ones(ord_.size, dtype=bool)
        sel[inds_locs - 1] = False
        ord_ = np.concatenate([ord_[sel], ord_[~sel]])
        obs = np.concatenate([obs[sel], obs[~sel]])
    return dict(U=U.tocsc(), latent=latent, ord=ord_, obs=obs, zero_nugg=zero_nugg,
                ord_pred=va["ord_pred"], ord_z=np.asarray(va["ord_z"]), cond_yz=va["cond_yz"],
                ic0=va.get("ic0", False), nfail=nfail)


# --------------------------------------------------------------------------------------------------
# vecchia_likelihood (R/vecchia_likelihood.R:14-99)
# --------------------------------------------------------------------------------------------------
def vecchia_loglik_numerator(z, vecchia_approx, covparms, nuggets, covmodel="matern", device=0):
    """quadform.num and logdet.num (vecchia_likelihood.R:74-76) without materialising U."""
    va = vecchia_approx
    n, size, latent, ord_, obs, nuggets, nuggets_all_ord, nuggets_ord = _prepare_nuggets(va, nuggets)
    if np.any(nuggets == 0):
        raise NotImplementedError("zero nuggets need the U trimming of createU.R:174-193; use createU")
    h = _handle_for(va, device)
    zord = np.asarray(z, dtype=np.float64)[np.asarray(va["ord_z"]) - 1]
    skip = n if va["cond_yz"] == "zy" else 0
    q, l, nfail = h.loglik_numerator(covmodel, covparms, nuggets_all_ord, nuggets_ord, zord, skip_rows=skip)
    return q, l, nfail


def _denominator(U_obj, z1):
    """vecchia_likelihood.R:85-91 via W = U.y U.y^T: logdet.denom = -log det W and
    quadform.denom = z2^T W^{-1} z2 (identical to the V = chol(rev W) formulation)."""
    latent = U_obj["latent"]
    U = U_obj["U"].tocsr()
    U_y = U[np.nonzero(latent)[0], :]
    z2 = np.asarray(U_y @ z1).ravel()
    W = (U_y @ U_y.T).tocsc()
    lu = spla.splu(W, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    logdetW = float(np.sum(np.log(np.abs(lu.L.diagonal()))) + np.sum(np.log(np.abs(lu.U.diagonal()))))
    quad = float(z2 @ lu.solve(z2))
    return -logdetW, quad


def vecchia_likelihood_U(z, U_obj):
    latent = U_obj["latent"]
    zord = np.asarray(z, dtype=np.float64)[U_obj["ord_z"] - 1]
    const = float((~latent).sum() * np.log(2 * np.pi))
    U = U_obj["U"].tocsr()
    z1 = np.asarray(U[np.nonzero(~latent)[0], :].T @ zord).ravel()
    quadform_num = float(np.sum(z1 ** 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        logdet_num = float(-2 * np.sum(np.log(U.diagonal())))
    if latent.sum() == 0:
        logdet_denom = quadform_denom = 0.0
    else:
        logdet_denom, quadform_denom = _denominator(U_obj, z1)
    neg2loglik = logdet_num - logdet_denom + quadform_num - quadform_denom + const
    return -neg2loglik / 2


def vecchia_likelihood(z, vecchia_approx, covparms, nuggets, covmodel="matern", device=0):
    va = vecchia_approx
    if va["cond_yz"] == "z" and bool(np.all(va["obs"])) and isinstance(covmodel, str) \
            and not np.any(np.asarray(nuggets) == 0):
        # standard Vecchia: numerator and denominator are per-row closed forms, all on the GPU
        n, size, latent, ord_, obs, nug, nuggets_all_ord, nuggets_ord = _prepare_nuggets(va, nuggets)
        h = _handle_for(va, device)
        zord = np.asarray(z, dtype=np.float64)[np.asarray(va["ord_z"]) - 1]
        return h.loglik_z(covmodel, covparms, nuggets_all_ord, nuggets_ord, zord)["loglik"]
    if vecchia_approx["cond_yz"] == "zy":
        import warnings
        warnings.warn("cond.yz='zy' will produce a poor likelihood approximation. Use 'SGV' instead.")
    U_obj = createU(vecchia_approx, covparms, nuggets, covmodel, device=device)
    return vecchia_likelihood_U(z, U_obj)
