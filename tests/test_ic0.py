"""ic0 / createUcppM / createUcpp (src/ic0.cpp), the incomplete-Cholesky branch of createU
(R/createU.R:89-106).

CPU: the oracle restatement against the reference's own known answer (tests/testthat/test-createL.r:43-45:
with the full lower-triangular pattern ic0 is the Cholesky factor, L L^T = Sigma to 1e-10 on the 20 x 20
grid with the exponential covariance, range 0.25), the library's host routine gpv_ic0 against the
restatement (bit-exact: same additions in the same order), argument errors as status codes.
GPU: gpv_createUcpp (covariance fill on the device + ic0) against the restated createUcpp."""
import numpy as np
import pytest

import oracle as O
import gpvecchia_b200 as G
from gpvecchia_b200 import _lib
from gpvecchia_b200 import harness as H


def _full_pattern(n):
    ptrs = np.concatenate([[0], np.cumsum(np.arange(1, n + 1))]).astype(np.float64)
    inds = np.concatenate([np.arange(i + 1) for i in range(n)]).astype(np.float64)
    return ptrs, inds


def _nn_pattern(locs, m):
    """Lower-triangular pattern of a nearest-neighbour Vecchia approximation, as createU.R:90-91 builds it
    from revNNarray: row i lists its earlier neighbours in ascending order... the reference takes them in
    revNNarray order (farthest first, self last); ic0's merge assumes ascending columns, which is what the
    MRA arrays provide, so the rows are sorted here."""
    NN = H.ordered_nn_kdtree(locs, m)          # (n, m+1): self first, then nearest earlier points, -1/NA padded
    rows = []
    for i in range(NN.shape[0]):
        r = np.asarray([v for v in NN[i] if v > 0], dtype=np.int64) - 1
        rows.append(np.sort(r))
    ptrs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.float64)
    inds = np.concatenate(rows).astype(np.float64)
    return ptrs, inds


def _mra_pattern(n, knots, block):
    """One-level multi-resolution pattern: the first `knots` rows are dense among themselves, every later row
    holds all knots and the earlier rows of its own block.  The exact Cholesky factor of the masked matrix
    has no fill outside this pattern and its Schur complements are conditional covariances, so ic0 cannot
    break down here (on a nearest-neighbour pattern it can: negative pivots -> NaN, as in the reference)."""
    rows = []
    for i in range(n):
        if i < knots:
            rows.append(np.arange(i + 1))
        else:
            b0 = knots + ((i - knots) // block) * block
            rows.append(np.concatenate([np.arange(knots), np.arange(b0, i + 1)]))
    ptrs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.float64)
    return ptrs, np.concatenate(rows).astype(np.float64)


def _dense_lower(ptrs, inds, vals, n):
    L = np.zeros((n, n))
    for i in range(n):
        a, b = int(ptrs[i]), int(ptrs[i + 1])
        L[i, inds[a:b].astype(int)] = vals[a:b]
    return L


def _grid_locs():
    g = np.linspace(0.0, 1.0, 20)
    X, Y = np.meshgrid(g, g)
    return np.column_stack([X.ravel(order="F"), Y.ravel(order="F")])    # expand.grid order


def test_oracle_ic0_reproduces_the_reference_known_answer():
    """test-createL.r:43-45: max |Sig0 - L00 L00^T| < 1e-10, Sig0 = exp(-rdist(locs) / 0.25)."""
    locs = _grid_locs()
    n = locs.shape[0]
    ptrs, inds = _full_pattern(n)
    vals = O.createUcpp(ptrs, inds, locs, [1.0, 0.25, 0.5])
    L = _dense_lower(ptrs, inds, vals, n)
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    Sig0 = np.exp(-D / 0.25)
    assert np.abs(Sig0 - L @ L.T).max() < 1e-10
    # createUcppM on the same covariances is the same factor
    cov = O.createUcpp(ptrs, inds, locs, [1.0, 0.25, 0.5], fill_only=True)
    assert np.array_equal(O.createUcppM(ptrs, inds, cov), vals)
    assert np.allclose(cov, Sig0[np.tril_indices(n)], rtol=1e-14, atol=0)


@pytest.mark.parametrize("pattern", ["full", "mra", "nn"])
def test_native_ic0_is_bit_identical_to_the_restatement(pattern):
    rng = np.random.default_rng(5)
    if pattern == "full":
        locs = _grid_locs()
        ptrs, inds = _full_pattern(locs.shape[0])
    elif pattern == "mra":
        locs = rng.random((1500, 2))
        ptrs, inds = _mra_pattern(1500, 40, 25)
    else:
        locs = rng.random((1500, 2))
        ptrs, inds = _nn_pattern(locs, 12)
    cov = O.createUcpp(ptrs, inds, locs, [1.3, 0.2, 1.5], fill_only=True)
    ref = O.ic0(ptrs, inds, cov)
    got = G.ic0(ptrs, inds, cov.copy())
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.array_equal(G.createUcppM(ptrs, inds, cov.copy()), ref, equal_nan=True)
    if pattern != "nn":
        assert np.isfinite(ref).all()
    if pattern == "mra":
        # the factor of the masked matrix: L L^T agrees with the covariance on the pattern
        n = locs.shape[0]
        L = _dense_lower(ptrs, inds, ref, n)
        S = _dense_lower(ptrs, inds, cov, n)
        LLt = L @ L.T
        mask = S != 0
        assert np.abs(LLt[mask] - S[mask]).max() < 1e-10
    # in place, like the reference
    buf = cov.copy()
    out = G.ic0(ptrs, inds, buf)
    assert out is buf and np.array_equal(buf, ref, equal_nan=True)


def test_ic0_nan_from_a_negative_pivot_propagates_like_the_reference():
    ptrs = np.array([0.0, 1.0, 3.0])
    inds = np.array([0.0, 0.0, 1.0])
    vals = np.array([1.0, 2.0, 1.0])          # second pivot 1 - 4 < 0 -> sqrt gives NaN (:56)
    ref = O.ic0(ptrs, inds, vals)
    got = G.ic0(ptrs, inds, vals.copy())
    assert np.isnan(ref[2]) and np.isnan(got[2]) and np.array_equal(got[:2], ref[:2])


def test_ic0_argument_errors_are_status_codes():
    ptrs = np.array([0.0, 1.0, 3.0])
    inds = np.array([0.0, 0.0, 1.0])
    vals = np.array([1.0, 0.5, 1.0])
    L = G.lib
    P = lambda a: a.ctypes.data
    assert L.gpv_ic0(2, None, P(inds), 3, P(vals)) == _lib.GPV_ERR_ARG
    assert L.gpv_ic0(2, P(ptrs), P(inds), 3, None) == _lib.GPV_ERR_ARG
    keep = vals.copy()
    bad = np.array([0.0, 1.0, 1.0])           # row 0 has an entry right of its diagonal (prints ERROR in the reference)
    bptrs = np.array([0.0, 2.0, 3.0])
    assert L.gpv_ic0(2, P(bptrs), P(bad), 3, P(vals)) == _lib.GPV_ERR_ARG and b"row 0" in L.gpv_last_error()
    assert np.array_equal(vals, keep)         # untouched
    assert L.gpv_ic0(2, P(np.array([0.0, 2.0, 1.0])), P(inds), 3, P(vals)) == _lib.GPV_ERR_ARG
    assert L.gpv_ic0(2, P(np.array([0.0, 1.5, 3.0])), P(inds), 3, P(vals)) == _lib.GPV_ERR_ARG
    assert L.gpv_ic0(2, P(ptrs), P(np.array([0.0, 0.5, 1.0])), 3, P(vals)) == _lib.GPV_ERR_ARG
    assert L.gpv_ic0(2, P(ptrs), P(inds), 2, P(vals)) == _lib.GPV_ERR_ARG      # ptrs[N] != nnz
    # a column whose own row is empty: the reference would read vals[-1]
    assert L.gpv_ic0(2, P(np.array([0.0, 0.0, 1.0])), P(np.array([0.0])), 1, P(vals)) == _lib.GPV_ERR_ARG
    assert L.gpv_createUcpp(2, 2, P(ptrs), P(inds), 3, None, None, P(vals), 0) == _lib.GPV_ERR_ARG
    locs = np.zeros(4)
    assert L.gpv_createUcpp(2, 2, P(ptrs), P(inds), 3, P(locs), P(np.array([1.0, 1.0, -1.0])), P(vals), 0) == _lib.GPV_ERR_ARG
    # empty problem
    assert L.gpv_ic0(0, P(np.array([0.0])), None, 0, None) == _lib.GPV_OK


@pytest.mark.gpu
@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5, 1.3])
def test_createUcpp_matches_the_restatement(nu):
    rng = np.random.default_rng(11)
    locs = rng.random((3000, 2))
    ptrs, inds = _mra_pattern(3000, 50, 30)
    covparams = [1.7, 0.15, nu]
    ref = O.createUcpp(ptrs, inds, locs, covparams)
    got = G.createUcpp(ptrs, inds, locs, covparams)
    assert np.isfinite(ref).all()
    # U-value bar of the north_star: 1e-10 relative, scaled by the row maximum
    n = locs.shape[0]
    for i in range(0, n, 7):
        a, b = int(ptrs[i]), int(ptrs[i + 1])
        assert np.abs(got[a:b] - ref[a:b]).max() <= 1e-10 * np.abs(ref[a:b]).max(), i
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.gpu
def test_createUcpp_full_pattern_known_answer_and_3d():
    """The reference's known answer through the device path: L L^T = Sigma (test-createL.r:43-45)."""
    locs = _grid_locs()
    n = locs.shape[0]
    ptrs, inds = _full_pattern(n)
    vals = G.createUcpp(ptrs, inds, locs, [1.0, 0.25, 0.5])
    L = _dense_lower(ptrs, inds, vals, n)
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    assert np.abs(np.exp(-D / 0.25) - L @ L.T).max() < 1e-10
    rng = np.random.default_rng(3)
    locs3 = rng.random((800, 3))
    p3, i3 = _mra_pattern(800, 30, 20)
    ref = O.createUcpp(p3, i3, locs3, [0.8, 0.3, 1.5])
    got = G.createUcpp(p3, i3, locs3, [0.8, 0.3, 1.5])
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
