"""Synthetic inputs for tests and bench.  NOT part of the reference path: orderings
and neighbour search are input producers that stay in the reference (vecchia_specify.R:100-178).

Locations: iid Uniform[0,1]^d from a counter-based generator (Philox, seed 20240601, stream =
config id); the index order is the ordering.  NNarray[k,] = (k, its m nearest among 1..k-1, nearest
first), NA-padded for the first m rows: the semantics of GpGp::find_ordered_nn as used at
R/vecchia_specify.R:159.  Conditioning layouts restate R/vecchia_specify.R:182-226.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check

from ._synth import (SEED, make_locs, make_data, make_nuggets, default_range, ordered_nn_kdtree, rev,  # noqa: F401
                     layout_yz, layout_zy)


def ordered_nn_gpu(locs, m, row_begin=0, row_end=None, device=0):
    """revNNarray (column-reversed, 1-based, 0-padded) for rows [row_begin,row_end) from the
    library's GPU grid search (gpv_harness_ordered_nn)."""
    locs = np.asarray(locs, dtype=np.float64)
    N, d = locs.shape
    if d > 3 or m > 63:
        raise ValueError("gpv_harness_ordered_nn supports d <= 3 and m <= 63; use ordered_nn_kdtree")
    row_end = N if row_end is None else row_end
    nrows = row_end - row_begin
    out = np.zeros((m + 1) * nrows, dtype=np.int32)
    lc = np.ascontiguousarray(locs.T).ravel()
    check(lib.gpv_harness_ordered_nn(N, d, m, lc.ctypes.data_as(C.c_void_p), row_begin, row_end,
                                     out.ctypes.data_as(C.c_void_p), device))
    return out.reshape(m + 1, nrows).T


def layout_zy_pred(locs_obs, locs_pred, m, device=0, use_gpu=True):
    """Response-first 'zy' layout WITH prediction locations in obs-then-pred ordering
    (vecchia_specify.R:120-149,191-224; BASELINE configs[4]): returns (locsord, NNarray, Cond, obs)
    with N = 2 n + n_pred rows (n dummy z rows with n0 = 1, then n + n_pred full rows).
    The ordered search for the prediction rows runs on the GPU harness, the unordered k-NN among the
    observations (FNN::get.knn) on a cKDTree."""
    from scipy.spatial import cKDTree
    locs_obs = np.asarray(locs_obs, dtype=np.float64)
    locs_pred = np.asarray(locs_pred, dtype=np.float64)
    n, n_p = locs_obs.shape[0], locs_pred.shape[0]
    locs_all = np.vstack([locs_obs, locs_pred])
    # ordered NN of the prediction rows among everything before them (:159)
    if use_gpu:
        rev = ordered_nn_gpu(locs_all, m, n, n + n_p, device=device).astype(np.int64)
        NN_pred = rev[:, ::-1]
    else:
        NN_pred = ordered_nn_kdtree(locs_all, m, n, n + n_p)
    tree = cKDTree(locs_obs)
    _, idx = tree.query(locs_obs, k=m, workers=-1)
    own = np.arange(n)[:, None]
    notself = idx != own
    # drop self (normally column 0); rows where self is missing from the k results are impossible
    # without exact duplicates, keep the first m-1 others
    order = np.argsort(~notself, axis=1, kind="stable")
    NNs = np.take_along_axis(idx, order, axis=1)[:, :m - 1].astype(np.int64) + 1
    prev = NNs < (own + 1)
    NNs[prev] += n
    NNarray_z = np.zeros((n, m + 1), dtype=np.int64)
    NNarray_z[:, 0] = np.arange(1, n + 1)
    NNarray_y = np.concatenate([own + 1 + n, own + 1, NNs], axis=1)
    NNarray_yp = NN_pred.copy()
    NNarray_yp[NNarray_yp != 0] += n
    NNarray = np.vstack([NNarray_z, NNarray_y, NNarray_yp])
    Cond = -np.ones(NNarray.shape, dtype=np.int8)
    nz = NNarray != 0
    Cond[nz] = (NNarray[nz] > n).astype(np.int8)
    Cond[:, 0] = 1
    obs = np.concatenate([np.ones(n, dtype=bool), np.zeros(n + n_p, dtype=bool)])
    return np.vstack([locs_obs, locs_all]), NNarray, Cond, obs


def whichCondOnLatent(NNarray, firstind_pred=None):
    """R/whichCondOnLatent.R through the library's host routine (gpv_whichCondOnLatent): NNarray is the
    un-reversed neighbour array (1-based, 0 / NA = missing); returns int8 1 / 0 / -1 (NA) like the
    restatement in oracle/vecchia_np.py."""
    NN = np.asarray(NNarray)
    n, p = NN.shape
    nn32 = np.asfortranarray(np.where(NN <= 0, np.iinfo(np.int32).min, NN).astype(np.int32))
    out = np.empty((n, p), dtype=np.int32, order="F")
    check(lib.gpv_whichCondOnLatent(nn32.ctypes.data_as(C.c_void_p), n, p, 0 if firstind_pred is None else int(firstind_pred),
                                    out.ctypes.data_as(C.c_void_p)))
    return np.where(out == np.iinfo(np.int32).min, -1, out).astype(np.int8)


def make_vecchia_approx(locsord, NNarray, Cond, obs, cond_yz, ord_=None, ord_z=None,
                        ord_pred="general", U_sparsity=None):
    """Assemble the vecchia.approx list (vecchia_specify.R:234-236) from harness-made inputs."""
    from .host import U_sparsity as us
    U_prep = (U_sparsity or us)(locsord, NNarray, obs, Cond)
    n = int(np.asarray(obs).sum())
    if ord_ is None:
        ord_ = np.arange(1, (n if cond_yz == "zy" else locsord.shape[0]) + 1)
    if ord_z is None:
        ord_z = np.arange(1, n + 1)
    return dict(locsord=locsord, obs=np.asarray(obs, dtype=bool), ord=ord_, ord_z=ord_z,
                ord_pred=ord_pred, U_prep=U_prep, cond_yz=cond_yz, ic0=False, conditioning="NN")
