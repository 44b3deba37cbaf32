"""Pure numpy / scipy part of the synthetic-input harness: no import of the C-ABI library, so
that bench.py's reference arm can load this file by path without mapping the product's .so.
gpvecchia_b200/harness.py re-exports everything here next to the helpers that do call the library."""
import numpy as np

SEED = 20240601


def make_locs(n, d, stream=0, seed=SEED):
    bg = np.random.Philox(key=seed, counter=[0, 0, 0, stream])
    return np.random.Generator(bg).random((n, d))


def make_data(n, stream=0, seed=SEED + 2):
    bg = np.random.Philox(key=seed, counter=[0, 0, 0, stream])
    return np.random.Generator(bg).standard_normal(n)


def make_nuggets(n, stream=0, seed=SEED + 1, lo=0.05, hi=0.15):
    bg = np.random.Philox(key=seed, counter=[0, 0, 0, stream])
    return lo + (hi - lo) * np.random.Generator(bg).random(n)


def default_range(n_obs, d):
    """range = 4 * n^(-1/d): a few neighbour spacings, keeps blocks well conditioned."""
    return 4.0 * float(n_obs) ** (-1.0 / d)


# --------------------------------------------------------------------------------------------------
# ordered nearest neighbours
# --------------------------------------------------------------------------------------------------
def ordered_nn_kdtree(locs, m, row_begin=0, row_end=None):
    """NNarray rows [row_begin,row_end) (1-based ids, 0 = NA) by doubling blocks of cKDTree
    queries.  Host harness; used for tests and as the cross-check of the GPU search."""
    from scipy.spatial import cKDTree
    locs = np.asarray(locs, dtype=np.float64)
    N = locs.shape[0]
    row_end = N if row_end is None else row_end
    NN = np.zeros((row_end - row_begin, m + 1), dtype=np.int64)
    NN[:, 0] = np.arange(row_begin + 1, row_end + 1)
    lo = 1
    while lo < N:
        hi = min(2 * lo, N)
        a, b = max(lo, row_begin), min(hi, row_end)
        if a < b:
            rows = np.arange(a, b)
            if hi <= 4096:
                D = np.sqrt(((locs[rows, None, :] - locs[None, :hi, :]) ** 2).sum(-1))
                D[np.arange(hi)[None, :] >= rows[:, None]] = np.inf
                order = np.argsort(D, axis=1, kind="stable")[:, :m]
                for r_i, r in enumerate(rows):
                    cnt = min(m, r)
                    NN[r - row_begin, 1:1 + cnt] = order[r_i, :cnt] + 1
            else:
                tree = cKDTree(locs[:hi])
                pending = rows
                k = min(hi, int(2.2 * m) + 12)
                while pending.size:
                    _, idx = tree.query(locs[pending], k=k, workers=-1)
                    mask = idx < pending[:, None]
                    rank = np.cumsum(mask, axis=1)
                    ok = rank[:, -1] >= m
                    sel = mask & (rank <= m)
                    good = pending[ok]
                    NN[good - row_begin, 1:] = idx[ok][sel[ok]].reshape(-1, m) + 1
                    pending = pending[~ok]
                    if k >= hi:
                        break
                    k = min(hi, 2 * k)
        lo = hi
    return NN


def rev(NNarray):
    return np.ascontiguousarray(np.asarray(NNarray)[:, ::-1])


# --------------------------------------------------------------------------------------------------
# conditioning layouts (vectorised restatement of vecchia_specify.R:182-226 for 'y', 'z', 'zy')
# --------------------------------------------------------------------------------------------------
def layout_yz(NNarray, cond_yz):
    """Cond (int8 1/0/-1) for cond.yz in {'y','z'} (:186-190)."""
    NNarray = np.asarray(NNarray)
    Cond = -np.ones(NNarray.shape, dtype=np.int8)
    Cond[NNarray != 0] = 1 if cond_yz == "y" else 0
    Cond[:, 0] = 1
    return Cond


def layout_zy(locsord, m, n):
    """Response-first 'zy' layout without prediction locations (:191-224): returns
    (locsord2, NNarray, Cond, obs) with N = 2n rows."""
    from scipy.spatial import cKDTree
    locs = np.asarray(locsord, dtype=np.float64)[:n]
    tree = cKDTree(locs)
    _, idx = tree.query(locs, k=m, workers=-1)          # self + (m-1) nearest
    idx = np.atleast_2d(idx)
    own = np.arange(n)[:, None]
    # drop self (normally column 0; be robust to duplicates)
    NNs = np.empty((n, m - 1), dtype=np.int64)
    for i in range(n):
        row = idx[i][idx[i] != i][:m - 1]
        NNs[i] = row
    NNs = NNs + 1
    prev = NNs < (own + 1)
    NNs[prev] += n
    NNarray_z = np.zeros((n, m + 1), dtype=np.int64)
    NNarray_z[:, 0] = np.arange(1, n + 1)
    NNarray_y = np.concatenate([own + 1 + n, own + 1, NNs], axis=1)
    NNarray = np.vstack([NNarray_z, NNarray_y])
    Cond = -np.ones(NNarray.shape, dtype=np.int8)
    nz = NNarray != 0
    Cond[nz] = (NNarray[nz] > n).astype(np.int8)
    Cond[:, 0] = 1
    obs = np.concatenate([np.ones(n, dtype=bool), np.zeros(n, dtype=bool)])
    return np.vstack([locs, locs]), NNarray, Cond, obs
