"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): nonzero pattern and indices bit-exact; U values within 1e-10
relative (row-max-scaled); likelihood terms within 1e-8 relative.
"""
import os

import numpy as np
import pytest

import gpvecchia_b200 as G
import oracle as O
from gpvecchia_b200 import harness as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NA_I32 = np.iinfo(np.int32).min
VAL_TOL = 1e-10
LL_TOL = 1e-8


def _pinned(n, src=None):
    """Page-locked float64 host array (numpy view of a pinned torch tensor): what the chunked copy pipeline needs."""
    import torch
    t = torch.empty(int(n), dtype=torch.float64).pin_memory()
    a = t.numpy()
    if src is not None:
        a[:] = src
    _PINNED_KEEPALIVE.append(t)
    return a


_PINNED_KEEPALIVE = []


def _rc_double(revCond):
    rc = revCond.astype(np.float64)
    rc[revCond < 0] = np.nan
    return rc


def _rowscaled_err(got, ref):
    scale = np.abs(ref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return float((np.abs(got - ref) / scale).max())


def _problem(n, m, d, cond_yz, stream):
    locs = H.make_locs(n, d, stream=stream)
    if cond_yz == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        return H.make_vecchia_approx(locs2, NN, Cond, obs, "zy")
    NN = H.ordered_nn_kdtree(locs, m)
    if cond_yz == "SGV":
        Cond = O.whichCondOnLatent(NN)
    else:
        Cond = H.layout_yz(NN, cond_yz)
    return H.make_vecchia_approx(locs, NN, Cond, np.ones(n, dtype=bool), cond_yz)


def _both(va, covType, covparms, nug_all, nug_obs):
    prep = va["U_prep"]
    n = int(va["obs"].sum())
    got = G.U_NZentries(1, n, va["locsord"], prep["revNNarray"], prep["revCond"], nug_all, nug_obs,
                        covType, covparms)
    ref = O.U_NZentries(O.max_threads(), n, va["locsord"], prep["revNNarray"], _rc_double(prep["revCond"]),
                        nug_all, nug_obs, covType, np.asarray(covparms, float))
    return got, ref


COVS = [("matern", [1.0, None, 0.5]), ("matern", [1.3, None, 1.5]), ("matern", [0.8, None, 2.5]),
        ("matern", [1.0, None, 0.8]), ("matern", [1.0, None, 1.3]), ("matern", [1.1, None, 3.2]),
        ("esqe", [0.7, None, 0.4, None])]


def _fill_range(cp, rng_):
    return [rng_ if v is None else v for v in cp]


@pytest.mark.parametrize("covType,cp", COVS)
def test_cfg2_shape_all_covariances(covType, cp):
    n, m, d = 4000, 30, 2
    va = _problem(n, m, d, "z", stream=2)
    cp = _fill_range(cp, H.default_range(n, d))
    nug = H.make_nuggets(n, stream=2)
    got, ref = _both(va, covType, cp, nug, nug)
    assert ref["nfail"] == 0 and got["nfail"] == 0
    assert np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL
    assert np.array_equal(got["Zentries"], ref["Zentries"]) or np.allclose(got["Zentries"], ref["Zentries"], rtol=2e-16, atol=0)


def test_cfg1_zy_layout_createU_and_likelihood():
    # BASELINE configs[0] at reduced n: vecchia_specify + vecchia_likelihood, 2-D, m=20, nu=1.5, zy
    n, m = 2000, 20
    va = _problem(n, m, 2, "zy", stream=1)
    cp = [1.0, H.default_range(n, 2), 1.5]
    z = H.make_data(n, stream=1)
    tau = H.make_nuggets(n, stream=1)
    Ug = G.createU(va, cp, tau)
    Uo = O.createU(va, cp, tau)
    A, B = Ug["U"].tocsc(), Uo["U"].tocsc()
    A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)   # pattern bit-exact
    assert np.array_equal(Ug["latent"], Uo["latent"]) and np.array_equal(Ug["obs"], Uo["obs"])
    # zy rows condition on latent y (no nugget) and hold a duplicated location (y_i, z_i): blocks are
    # ill conditioned, so two correct fp64 implementations differ by cond*eps.  Arbitrate with the
    # __float128 oracle: the CUDA path must be as close to it as the fp64 restatement.
    Uq = O.createU(va, cp, tau, mode=2)["U"].tocsc()
    Uq.sort_indices()
    colmax = np.maximum.reduceat(np.abs(Uq.data), Uq.indptr[:-1])
    scale = np.repeat(colmax, np.diff(Uq.indptr))
    err_gpu = (np.abs(A.data - Uq.data) / scale).max()
    err_ref = (np.abs(B.data - Uq.data) / scale).max()
    assert err_gpu < max(VAL_TOL, 3 * err_ref), (err_gpu, err_ref)
    assert (np.abs(A.data - B.data) / scale).max() < 1e-8
    q, l, nf = G.vecchia_loglik_numerator(z, va, cp, tau)
    qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
    assert nf == 0 and abs(q - qr) <= LL_TOL * abs(qr) and abs(l - lr) <= LL_TOL * abs(lr)
    with pytest.warns(UserWarning):
        ll = G.vecchia_likelihood(z, va, cp, tau)
    llr = O.vecchia_likelihood_U(z, Uo)
    assert abs(ll - llr) <= LL_TOL * abs(llr)


@pytest.mark.parametrize("cond_yz", ["y", "z", "SGV"])
def test_full_loglik_against_oracle_and_exact(cond_yz):
    n, m = 600, 15
    va = _problem(n, m, 2, cond_yz, stream=6)
    cp = [1.2, 0.15, 1.5]
    z = H.make_data(n, stream=6)
    ll = G.vecchia_likelihood(z, va, cp, 0.1)
    llr = O.vecchia_likelihood(z, va, cp, 0.1)
    assert abs(ll - llr) <= LL_TOL * abs(llr)
    q, l, _ = G.vecchia_loglik_numerator(z, va, cp, 0.1)
    qr, lr, _ = O.loglik_numerator_from_U(z, O.createU(va, cp, 0.1))
    assert abs(q - qr) <= LL_TOL * abs(qr) and abs(l - lr) <= LL_TOL * abs(lr)


def test_known_answer_full_conditioning_exact_density():
    # m = n-1 => exact Gaussian log-density (vignette :128-139); n-1 = 30 keeps p = 31
    n = 31
    locs = H.make_locs(n, 2, stream=7)
    z = H.make_data(n, stream=7)
    for cyz in ("y", "z", "SGV"):
        NN = H.ordered_nn_kdtree(locs, n - 1)
        Cond = O.whichCondOnLatent(NN) if cyz == "SGV" else H.layout_yz(NN, cyz)
        va = H.make_vecchia_approx(locs, NN, Cond, np.ones(n, dtype=bool), cyz)
        for cp in ([1.3, 0.25, 1.5], [1.3, 0.25, 0.8], [1.0, 0.3, 2.5]):
            ll = G.vecchia_likelihood(z, va, cp, 0.2)
            ex = O.exact_loglik(z, locs, cp, 0.2)
            assert abs(ll - ex) <= LL_TOL * abs(ex), (cyz, cp)


def test_quad_precision_fixture():
    f = np.load(os.path.join(GOLD, "u_small_quad.npz"))
    n = f["locs"].shape[0]
    for tag, ct in [("m05", "matern"), ("m15", "matern"), ("m25", "matern"), ("g08", "matern"),
                    ("g13", "matern"), ("esqe", "esqe")]:
        got = G.U_NZentries(1, n, f["locs"], f["revNNarray"], f["revCond"], f["nuggets"], f["nuggets"],
                            ct, f["cp_" + tag])
        gold = f["L_" + tag]
        assert np.array_equal(got["Lentries"] == 0, gold == 0), tag
        assert _rowscaled_err(got["Lentries"], gold) < VAL_TOL, tag


@pytest.mark.parametrize("m", [1, 2, 3, 7, 10, 12, 15, 20, 25, 30, 31])
def test_every_instantiated_set_size(m):
    n = 1500
    va = _problem(n, m, 2, "SGV" if m <= 10 else "z", stream=10 + m)
    cp = [1.0, H.default_range(n, 2), 1.5]
    nug = np.full(n, 0.1)
    got, ref = _both(va, "matern", cp, nug, nug)
    assert np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5])
def test_spatial_dimensions(d):
    n, m = 1200, 10
    va = _problem(n, m, d, "z", stream=40 + d)
    cp = [1.0, H.default_range(n, d), 2.5]
    nug = np.full(n, 0.05)
    got, ref = _both(va, "matern", cp, nug, nug)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL
    cp = [0.6, H.default_range(n, d), 0.5, 0.7 * H.default_range(n, d)]
    got, ref = _both(va, "esqe", cp, nug, nug)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL


def test_edge_duplicates_zero_inf_and_negative_nuggets():
    n, m = 400, 8
    locs = H.make_locs(n, 2, stream=50)
    locs[100:110] = locs[90:100]              # exact duplicates: D == 0 off the diagonal
    NN = H.ordered_nn_kdtree(locs, m)
    va = H.make_vecchia_approx(locs, NN, H.layout_yz(NN, "z"), np.ones(n, dtype=bool), "z")
    cp = [1.0, 0.2, 1.5]
    nug = np.full(n, 0.1)
    nug[5] = 0.0                               # zero nugget (createU.R:83-86 territory)
    nug[7] = np.inf                            # VL missing data (vecchia_laplace_NR.R:108)
    nug[300] = -40.0                           # indefinite blocks -> zero rows (U_NZentries.cpp:64-66)
    got, _ = _both(va, "matern", cp, nug, np.abs(nug))
    # Row 8 has an Inf self-nugget times (1 - revCond) = 0 -> NaN diagonal (U_NZentries.cpp:47).
    # Reference LAPACK's dpotrf/dpotf2 tests `ajj <= 0 || disnan(ajj)` and fails, chol() throws and
    # the row stays zero; the OpenBLAS build inside scipy skips the NaN test and returns a NaN row.
    # The CUDA path follows published LAPACK, i.e. oracle mode 1 (textbook dpotf2).
    prep = va["U_prep"]
    ref = O.U_NZentries(2, n, va["locsord"], prep["revNNarray"], _rc_double(prep["revCond"]), nug, np.abs(nug),
                        "matern", np.array(cp), mode=1)
    assert got["nfail"] == ref["nfail"] > 0
    bad = np.nonzero(np.all(ref["Lentries"] == 0, axis=1))[0]
    assert got["first_fail"] == bad.min()
    assert np.all(got["Lentries"][bad] == 0)
    ok = np.ones(n, dtype=bool); ok[bad] = False
    assert np.array_equal(np.isnan(got["Lentries"]), np.isnan(ref["Lentries"]))
    fin = np.isfinite(ref["Lentries"]).all(axis=1) & ok
    assert _rowscaled_err(got["Lentries"][fin], ref["Lentries"][fin]) < VAL_TOL
    # Z entries: 0 nugget -> -/+Inf, Inf nugget -> -/+0 (U_NZentries.cpp:112-113)
    assert np.array_equal(got["Zentries"], ref["Zentries"])


@pytest.mark.parametrize("m", [8, 30])
@pytest.mark.parametrize("cov,cp", [("matern", [1.0, 0.2, 0.5]), ("matern", [1.0, 0.2, 1.5]), ("matern", [1.0, 0.2, 2.5]),
                                    ("matern", [1.0, 0.2, 0.8]), ("esqe", [0.7, 0.2, 0.3, 0.1])])
def test_nan_coordinate_fails_exactly_the_rows_that_see_it(m, cov, cp):
    # a NaN coordinate makes every distance to that point NaN, the covariance NaN (no `dist == 0`
    # branch catches it) and dpotrf's isnan test fails the block (U_NZentries.cpp:60-66): all rows
    # whose conditioning set holds the point stay zero, every other row is unaffected.  Guards the
    # integer clamp inside exp (u_kernels.cuh), which must let a NaN argument through.
    n = 600
    locs = H.make_locs(n, 2, stream=53)
    NN = H.ordered_nn_kdtree(locs, m)
    locs[37, 1] = np.nan
    va = H.make_vecchia_approx(locs, NN, H.layout_yz(NN, "z"), np.ones(n, dtype=bool), "z")
    prep = va["U_prep"]
    nug = np.full(n, 0.1)
    got = G.U_NZentries(1, n, va["locsord"], prep["revNNarray"], prep["revCond"], nug, nug, cov, cp)
    sees = np.any(prep["revNNarray"] == 38, axis=1)
    assert got["nfail"] == int(sees.sum()) and got["first_fail"] == int(np.nonzero(sees)[0].min())
    assert np.all(got["Lentries"][sees] == 0)
    clean = locs.copy(); clean[37, 1] = 0.5
    ref = O.U_NZentries(2, n, clean, prep["revNNarray"], _rc_double(prep["revCond"]), nug, nug, cov, np.array(cp))
    assert _rowscaled_err(got["Lentries"][~sees], ref["Lentries"][~sees]) < VAL_TOL


@pytest.mark.parametrize("cond_yz", ["z", "y", "SGV", "zy"])
def test_native_sparsity_and_csc_output(cond_yz):
    # U_sparsity's triplet arrays (R/U_sparsity.R:36-73) and the dgCMatrix slots of
    # sparseMatrix(i, j, x) (R/createU.R:161) from the library, bit-exact against the restated R loops
    n, m = 500, 9
    va = _problem(n, m, 2, cond_yz, stream=60)
    prep = va["U_prep"]
    N = va["locsord"].shape[0]
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
        ci, rp = h.u_sparsity()
        assert ci.dtype == np.int32 and np.array_equal(ci, prep["colindices"]) and np.array_equal(rp, prep["rowpointers"])
        ncols, nnz, size = h.csc_dims()
        assert size == prep["size"] == ncols and nnz == prep["colindices"].size
        colptr, rowidx = h.csc_pattern()
        import scipy.sparse as sp
        pat = sp.coo_matrix((np.ones(nnz), (prep["colindices"] - 1, prep["rowpointers"] - 1)), shape=(size, size)).tocsc()
        pat.sort_indices()
        assert np.array_equal(colptr, pat.indptr) and np.array_equal(rowidx, pat.indices)
        # shards: the slices of consecutive row ranges concatenate to the full arrays
        cuts = [0, 1, 130, 131, N]
        cps, ris, base = [], [], 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"], row_begin=a, row_end=b) as hs:
                cp_s, ri_s = hs.csc_pattern()
                cps.append(cp_s[:-1].astype(np.int64) + base); base += int(cp_s[-1]); ris.append(ri_s)
        assert np.array_equal(np.concatenate(cps + [[base]]), colptr) and np.array_equal(np.concatenate(ris), rowidx)
    tau = H.make_nuggets(n, stream=60)
    cp = [1.3, 0.15, 1.5]
    A = G.createU(va, cp, tau, assemble="csc")["U"]
    B = G.createU(va, cp, tau, assemble="triplet")["U"]
    A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    Cm = O.createU(va, cp, tau)["U"].tocsc(); Cm.sort_indices()
    assert np.array_equal(A.indptr, Cm.indptr) and np.array_equal(A.indices, Cm.indices)
    # values: column-scaled (a column of U is one conditioning set); latent conditioning without nugget
    # is ill conditioned at this range, so the bound is cond * eps, not VAL_TOL
    colmax = np.maximum.reduceat(np.abs(Cm.data), Cm.indptr[:-1])
    assert (np.abs(A.data - Cm.data) / np.repeat(colmax, np.diff(Cm.indptr))).max() < 1e-8


def test_csc_values_at_scale_match_packed_order():
    # n = 2e5, m = 30: x[perm] == packed values exactly, the permutation being the one the pattern implies
    n, m = 200_000, 30
    locs = H.make_locs(n, 2, stream=61)
    revNN = H.ordered_nn_gpu(locs, m)
    revCond = np.zeros(revNN.shape, dtype=np.int32)
    revCond[revNN == 0] = np.iinfo(np.int32).min
    revCond[:, -1] = 1
    nug = H.make_nuggets(n, stream=61)
    cp = [1.0, H.default_range(n, 2), 1.5]
    with G.UHandle(locs, revNN, revCond, obs=np.ones(n, dtype=np.int32)) as h:
        packed, _, _ = h.values_packed("matern", cp, nug, nug)
        x, nf, _ = h.values_csc("matern", cp, nug, nug)
        ci, rp = h.u_sparsity()
        colptr, rowidx = h.csc_pattern()
    assert nf == 0 and x.size == packed.size
    import scipy.sparse as sp
    size = 2 * n
    M = sp.csc_matrix((x, rowidx, colptr), shape=(size, size))
    T = sp.coo_matrix((packed, (ci - 1, rp - 1)), shape=(size, size)).tocsc()
    T.sort_indices()
    assert np.array_equal(M.indptr, T.indptr) and np.array_equal(M.indices, T.indices) and np.array_equal(M.data, T.data)


@pytest.mark.parametrize("m,cond_yz", [(8, "z"), (20, "SGV"), (40, "z"), (9, "zy")])
def test_covmodel_matrix_branch(m, cond_yz):
    # U_NZentries_mat (src/U_NZentries.cpp:126-197, createU.R:149-151): covmat = covVals(inds, inds), no
    # nugget added, revCond not read.  Values and packed order vs the restatement; one indefinite
    # block (a row/column of covVals scaled to break positive definiteness) leaves exactly its rows zero.
    n = 260
    va = _problem(n, m, 2, cond_yz, stream=70 + m)
    prep = va["U_prep"]
    locs = va["locsord"]
    N = locs.shape[0]
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    covVals = O.MaternFun(D, np.array([1.2, 0.25, 1.5])) + 0.05 * np.eye(N)
    nobs = int(va["obs"].sum())
    tau = H.make_nuggets(nobs, stream=70)
    with G.UHandle(locs, prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
        got = h.U_NZentries_mat(covVals, tau)
        ref = O.U_NZentries_mat(nobs, prep["revNNarray"], covVals, tau)
        assert got["nfail"] == ref["nfail"] == 0
        assert np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
        assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL
        assert np.array_equal(got["Zentries"], ref["Zentries"])
        packed, _, _ = h.values_packed_mat(covVals, tau)
        not_na = (prep["revNNarray"][:, ::-1] != 0).ravel() & (prep["revNNarray"][:, ::-1] != NA_I32).ravel()
        assert np.array_equal(packed, np.concatenate([got["Lentries"].ravel()[not_na], got["Zentries"]]))
        bad = covVals.copy()
        bad[N - 3, N - 3] = -1.0
        g2, r2 = h.U_NZentries_mat(bad, tau), O.U_NZentries_mat(nobs, prep["revNNarray"], bad, tau)
        assert g2["nfail"] == r2["nfail"] > 0
        assert np.array_equal(np.all(g2["Lentries"] == 0, axis=1), np.all(r2["Lentries"] == 0, axis=1))
    # through createU: same sparse matrix as the triplet assembly of the restated values
    if cond_yz != "zy":
        Ug = G.createU(va, [1.2, 0.25, 1.5], tau, covmodel=covVals)["U"].tocsc()
        Ug.sort_indices()
        import scipy.sparse as sp
        want = np.concatenate([ref["Lentries"].ravel()[not_na], ref["Zentries"]])
        Uo = sp.coo_matrix((want, (prep["colindices"] - 1, prep["rowpointers"] - 1)), shape=Ug.shape).tocsc()
        Uo.sort_indices()
        assert np.array_equal(Ug.indptr, Uo.indptr) and np.array_equal(Ug.indices, Uo.indices)
        colmax = np.maximum.reduceat(np.abs(Uo.data), Uo.indptr[:-1])
        assert (np.abs(Ug.data - Uo.data) / np.repeat(colmax, np.diff(Uo.indptr))).max() < VAL_TOL


def test_likelihood_reuses_resident_data_and_scalar_nugget():
    # estimation loop (R/vecchia_wrappers.R:72-93): z fixed, nugget one scalar -> None arguments reuse what
    # is resident on the handle; gpv_set_scalar_nugget builds nuggets.all.ord / nuggets.ord on the device
    n, m = 3000, 15
    for cond_yz in ("z", "zy"):
        va = _problem(n, m, 2, cond_yz, stream=80)
        prep = va["U_prep"]
        N = va["locsord"].shape[0]
        z = H.make_data(n, stream=80)
        nug_all = np.concatenate([np.full(n, 0.07), np.zeros(N - n)])
        tau = np.full(n, 0.07)
        skip = n if cond_yz == "zy" else 0
        with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
            with pytest.raises(G.GpvError):
                h.loglik_numerator("matern", [1.0, 0.1, 1.5], None, None, None, skip_rows=skip)
            want = [h.loglik_numerator("matern", [1.0, r, 1.5], nug_all, tau, z, skip_rows=skip) for r in (0.1, 0.2)]
            got = [h.loglik_numerator("matern", [1.0, r, 1.5], None, None, None, skip_rows=skip) for r in (0.1, 0.2)]
            assert got == want                                   # same kernel, same device data: bit-identical
            h.set_scalar_nugget(0.07)
            assert h.loglik_numerator("matern", [1.0, 0.2, 1.5], None, None, None, skip_rows=skip) == want[1]
            h.set_scalar_nugget(0.11)
            a = h.loglik_numerator("matern", [1.0, 0.2, 1.5], None, None, None, skip_rows=skip)
            b = h.loglik_numerator("matern", [1.0, 0.2, 1.5], nug_all / 0.07 * 0.11, tau / 0.07 * 0.11, z, skip_rows=skip)
            assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-12 * abs(b[1])
            if cond_yz == "z":
                r1 = h.loglik_z("matern", [1.0, 0.2, 1.5], None, None, None)
                r2 = h.loglik_z("matern", [1.0, 0.2, 1.5], nug_all / 0.07 * 0.11, tau / 0.07 * 0.11, z)
                assert abs(r1["loglik"] - r2["loglik"]) <= 1e-12 * abs(r2["loglik"])
            # U values with the resident scalar nugget (no per-call nugget upload): the same bits as with the vectors,
            # and the vectors stay resident for the next call
            want_p = h.values_packed("matern", [1.0, 0.2, 1.5], nug_all / 0.07 * 0.11, tau / 0.07 * 0.11)
            h.set_scalar_nugget(0.11)
            for _ in range(2):
                got_p = h.values_packed("matern", [1.0, 0.2, 1.5], None, None)
                assert got_p[1:] == want_p[1:] and np.array_equal(got_p[0], want_p[0])
            want_c = h.values_csc("matern", [1.0, 0.2, 1.5], nug_all / 0.07 * 0.11, tau / 0.07 * 0.11)
            h.set_scalar_nugget(0.11)
            got_c = h.values_csc("matern", [1.0, 0.2, 1.5], None, None)
            assert np.array_equal(got_c[0], want_c[0])
            assert h.loglik_numerator("matern", [1.0, 0.2, 1.5], None, None, None, skip_rows=skip) == a
            # a U-values call that brings nugget vectors overwrites the buffer: the next reuse must be refused, not
            # silently wrong
            h.values_packed("matern", [1.0, 0.2, 1.5], nug_all, tau)
            with pytest.raises(G.GpvError):
                h.loglik_numerator("matern", [1.0, 0.2, 1.5], None, None, None, skip_rows=skip)
            with pytest.raises(G.GpvError):
                h.values_packed("matern", [1.0, 0.2, 1.5], None, None)


def test_zero_nugget_createU_trimming():
    n, m = 300, 6
    va = _problem(n, m, 2, "SGV", stream=51)
    tau = np.full(n, 0.1)
    tau[[3, 50, 299]] = 0.0
    cp = [1.0, 0.2, 1.5]
    Ug, Uo = G.createU(va, cp, tau), O.createU(va, cp, tau)
    A, B = Ug["U"].tocsc(), Uo["U"].tocsc()
    A.sort_indices(); B.sort_indices()
    assert A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.allclose(A.data, B.data, rtol=1e-9, atol=0)
    for k in ("inds_U", "inds_z", "inds_locs"):
        assert np.array_equal(Ug["zero_nugg"][k], Uo["zero_nugg"][k])
    assert np.array_equal(Ug["latent"], Uo["latent"]) and np.array_equal(Ug["ord"], Uo["ord"])


def test_missing_entries_anywhere_are_compacted_like_find():
    # `inds.elem(find(inds))` (U_NZentries.cpp:44) compacts zeros wherever they are; revCond is read
    # from the LAST n0 columns (:47)
    n, m = 200, 6
    va = _problem(n, m, 2, "z", stream=52)
    prep = va["U_prep"]
    rnn = prep["revNNarray"].copy()
    rnn[50:120:7, 2] = 0
    rnn[60:130:9, 0] = 0
    nug = np.full(n, 0.1)
    cp = [1.0, 0.3, 0.5]
    got = G.U_NZentries(1, n, va["locsord"], rnn, prep["revCond"], nug, nug, "matern", cp)
    ref = O.U_NZentries(1, n, va["locsord"], rnn, _rc_double(prep["revCond"]), nug, nug, "matern", np.array(cp))
    assert np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL


def test_handle_reuse_packed_order_and_row_shards():
    n, m = 3000, 12
    va = _problem(n, m, 2, "z", stream=53)
    prep = va["U_prep"]
    nug = H.make_nuggets(n, stream=53)
    obs = np.ones(n, dtype=bool)
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=obs) as h:
        outs = {}
        for cp in ([1.0, 0.05, 1.5], [2.0, 0.08, 0.5], [1.0, 0.05, 1.7]):
            r = h.U_NZentries("matern", cp, nug, nug)
            ref = O.U_NZentries(2, n, va["locsord"], prep["revNNarray"], _rc_double(prep["revCond"]), nug, nug,
                                "matern", np.array(cp))
            assert _rowscaled_err(r["Lentries"], ref["Lentries"]) < VAL_TOL
            packed, nf, _ = h.values_packed("matern", cp, nug, nug)
            not_na = (prep["revNNarray"][:, ::-1] != 0).ravel()
            want = np.concatenate([r["Lentries"].ravel()[not_na], r["Zentries"]])   # createU.R:158-160
            assert np.array_equal(packed, want)
            outs[tuple(cp)] = r["Lentries"]
        assert h.last_kernel_name().startswith(("u_sets<", "u_quad<")) and h.last_kernel_ms() > 0
        full = outs[(1.0, 0.05, 1.5)]
    cuts = [0, 1, 17, 1000, 1000, 3000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=obs, row_begin=a, row_end=b) as hs:
            r = hs.U_NZentries("matern", [1.0, 0.05, 1.5], nug, nug)
            assert r["Lentries"].shape == (b - a, m + 1)
            assert np.array_equal(r["Lentries"], full[a:b])      # same kernel, same rows: bit-identical


def test_sharded_loglik_partials_sum_to_the_whole():
    n, m = 4000, 10
    va = _problem(n, m, 2, "z", stream=54)
    prep = va["U_prep"]
    nug = H.make_nuggets(n, stream=54)
    z = H.make_data(n, stream=54)
    obs = np.ones(n, dtype=bool)
    cp = [1.0, 0.04, 0.8]
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=obs) as h:
        q, l, nf = h.loglik_numerator("matern", cp, nug, nug, z)
    qs = ls = 0.0
    for a, b in ((0, 1300), (1300, 2600), (2600, 4000)):
        with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=obs, row_begin=a, row_end=b) as hs:
            qq, lll, _ = hs.loglik_numerator("matern", cp, nug, nug, z)
            qs += qq; ls += lll
    assert abs(qs - q) <= 1e-12 * abs(q) and abs(ls - l) <= 1e-12 * abs(l)
    Uo = O.createU(va, cp, nug)
    qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
    assert abs(q - qr) <= LL_TOL * abs(qr) and abs(l - lr) <= LL_TOL * abs(lr)


def test_covariance_functions_alone():
    # the reference's test-MaternFun.r, against the device functions
    locs = H.make_locs(100, 2, stream=55)
    D = np.sqrt(((locs[:, None] - locs[None]) ** 2).sum(-1))
    for nu in (0.5, 1.5, 2.5):
        assert np.abs(G.MaternFun(D, [1.0, 0.2, nu]) - O.MaternFun(D, [1.0, 0.2, nu])).sum() < 1e-10
    for nu in (0.3, 0.8, 1.3, 4.1):
        a, b = G.MaternFun(D, [1.0, 0.2, nu]), O.MaternFun(D, [1.0, 0.2, nu])
        assert np.abs(a - b).sum() < 1e-10 and (np.abs(a - b) / b).max() < 1e-12
    cp = [0.7, 0.2, 0.4, 0.1]
    assert np.abs(G.EsqeFun(D, cp) - O.EsqeFun(D, cp)).sum() < 1e-10
    assert G.MaternFun(np.zeros(4), [2.5, 0.2, 0.8]).tolist() == [2.5] * 4


def test_scale_property_at_larger_n():
    # size-independent property: cov -> c*cov (sig2 and nuggets scaled by c) => U -> U / sqrt(c)
    n, m = 60000, 30
    locs = H.make_locs(n, 2, stream=56)
    NN = H.ordered_nn_kdtree(locs, m)
    revNN = H.rev(NN)
    revCond = H.rev(H.layout_yz(NN, "z"))
    nug = H.make_nuggets(n, stream=56)
    c = 4.0
    with G.UHandle(locs, revNN, revCond) as h:
        rng_ = H.default_range(n, 2)
        a = h.U_NZentries("matern", [1.0, rng_, 1.5], nug, nug)["Lentries"]
        b = h.U_NZentries("matern", [c, rng_, 1.5], c * nug, c * nug)["Lentries"]
    assert _rowscaled_err(b * np.sqrt(c), a) < 1e-12
    # diagonal of U is positive, rows beyond n0 are zero
    assert np.all(a[:, -1][m:] > 0)


@pytest.mark.parametrize("d,m,n", [(2, 30, 50000), (3, 12, 20000), (1, 5, 5000), (2, 3, 1000)])
def test_harness_gpu_ordered_nn_matches_host_search(d, m, n):
    # harness, not the reference path: the GPU grid search must produce the arrays the host
    # (cKDTree / brute force) search produces, bit for bit, including a row shard
    locs = H.make_locs(n, d, stream=60 + d)
    ref = H.rev(H.ordered_nn_kdtree(locs, m))
    got = H.ordered_nn_gpu(locs, m)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    a, b = n // 3, min(n, n // 3 + 777)
    assert np.array_equal(H.ordered_nn_gpu(locs, m, a, b), ref[a:b])


def test_cfg4_shape_m40_3d_esqe():
    # BASELINE configs[3] at reduced n: 3-D locations, m = 40 (p = 41 > 32: one set per warp), esqe
    n, m, d = 3000, 40, 3
    va = _problem(n, m, d, "z", stream=80)
    rng_ = H.default_range(n, d)
    cp = [1.0, rng_, 0.5, rng_]
    nug = H.make_nuggets(n, stream=80)
    got, ref = _both(va, "esqe", cp, nug, nug)
    assert got["nfail"] == 0 and np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL
    z = H.make_data(n, stream=80)
    q, l, _ = G.vecchia_loglik_numerator(z, va, cp, nug, covmodel="esqe")
    qr, lr, _ = O.loglik_numerator_from_U(z, O.createU(va, cp, nug, covmodel="esqe"))
    assert abs(q - qr) <= LL_TOL * abs(qr) and abs(l - lr) <= LL_TOL * abs(lr)


@pytest.mark.parametrize("m", [33, 40, 45, 50, 63])
def test_large_set_sizes(m):
    n = 1200
    va = _problem(n, m, 2, "z", stream=90 + m)
    cp = [1.0, H.default_range(n, 2), 0.5]
    nug = np.full(n, 0.1)
    got, ref = _both(va, "matern", cp, nug, nug)
    assert np.array_equal(got["Lentries"] == 0, ref["Lentries"] == 0)
    assert _rowscaled_err(got["Lentries"], ref["Lentries"]) < VAL_TOL


def test_cfg5_obs_pred_joint_ordering_zy():
    # BASELINE configs[4] at reduced n: obs-then-pred ordering, default cond.yz = 'zy' with
    # prediction locations (vecchia_specify.R:92-96,191-224): N = 2 n_obs + n_pred rows
    n_obs, n_pred, m = 700, 180, 30
    locs = H.make_locs(n_obs, 2, stream=85)
    locs_pred = H.make_locs(n_pred, 2, stream=86)
    va = O.vecchia_specify(locs, m, cond_yz="zy", locs_pred=locs_pred)
    assert va["locsord"].shape[0] == 2 * n_obs + n_pred
    cp = [1.0, H.default_range(n_obs, 2), 1.5]
    tau = H.make_nuggets(n_obs, stream=85)
    Ug, Uo, Uq = G.createU(va, cp, tau), O.createU(va, cp, tau), O.createU(va, cp, tau, mode=2)
    A, B, Q = Ug["U"].tocsc(), Uo["U"].tocsc(), Uq["U"].tocsc()
    for M in (A, B, Q):
        M.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    colmax = np.maximum.reduceat(np.abs(Q.data), Q.indptr[:-1])
    scale = np.repeat(colmax, np.diff(Q.indptr))
    err_gpu = (np.abs(A.data - Q.data) / scale).max()
    err_ref = (np.abs(B.data - Q.data) / scale).max()
    assert err_gpu < max(VAL_TOL, 3 * err_ref), (err_gpu, err_ref)
    z = H.make_data(n_obs, stream=85)
    q, l, _ = G.vecchia_loglik_numerator(z, va, cp, tau)
    qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
    assert abs(q - qr) <= LL_TOL * abs(qr) and abs(l - lr) <= LL_TOL * abs(lr)


def test_shard_arrays_entry_point_and_kernel_timing():
    # gpv_create_shard: a rank hands over only its own rows of revNNarray / revCond
    n, m = 2500, 9
    va = _problem(n, m, 2, "z", stream=95)
    prep = va["U_prep"]
    nug = H.make_nuggets(n, stream=95)
    z = H.make_data(n, stream=95)
    obs = np.ones(n, dtype=bool)
    cp = [1.0, 0.05, 2.5]
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=obs) as h:
        full = h.U_NZentries("matern", cp, nug, nug)["Lentries"]
        q, l, _ = h.loglik_numerator("matern", cp, nug, nug, z)
        cnt, tot = h.kernel_time_stats()
        assert cnt == 2 and tot > 0
    a, b = 700, 1900
    with G.UHandle(va["locsord"], prep["revNNarray"][a:b], prep["revCond"][a:b], obs=obs, row_begin=a, row_end=b) as hs:
        part = hs.U_NZentries("matern", cp, nug, nug)["Lentries"]
        assert np.array_equal(part, full[a:b])
        packed, _, _ = hs.values_packed("matern", cp, nug, nug, zentries_tail=False)
        assert packed.size == (b - a) * (m + 1) and np.array_equal(packed, full[a:b].ravel())
    with pytest.raises(ValueError):
        G.UHandle(va["locsord"], prep["revNNarray"][a:b], prep["revCond"][a:b], obs=obs, row_begin=a, row_end=b + 5)


@pytest.mark.parametrize("layout", ["z", "zy"])
def test_chunked_packed_pipeline_matches_single_launch(layout):
    # gpv_u_values_packed overlaps kernel chunks with D2H copies once there are >= 65536 sets;
    # the result must be bit-identical to the unchunked row-major output, also with the
    # n0 <= 1 rows of a zy layout interleaved (set list + closed-form kernel)
    n, m = 70000, 10
    locs = H.make_locs(n, 2, stream=97)
    if layout == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        nug_all = np.concatenate([H.make_nuggets(n, stream=97), np.zeros(n)])
    else:
        locs2, NN = locs, H.ordered_nn_kdtree(locs, m)
        Cond, obs = H.layout_yz(NN, "z"), np.ones(n, dtype=bool)
        nug_all = H.make_nuggets(n, stream=97)
    revNN, revCond = H.rev(NN), H.rev(Cond)
    tau = nug_all[:n]
    cp = [1.0, H.default_range(n, 2), 1.5]
    with G.UHandle(locs2, revNN, revCond, obs=obs) as h:
        r = h.U_NZentries("matern", cp, nug_all, tau)
        packed, nf, _ = h.values_packed("matern", cp, nug_all, tau)          # pageable destination: one launch, one copy
        out = _pinned(packed.size)                                           # page-locked: the chunked pipeline, with the
        out[:] = np.nan                                                      # nugget upload staged chunk by chunk
        packed2, nf2, _ = h.values_packed("matern", cp, nug_all, tau, out=out)
        _, nnz, _ = h.csc_dims()
        x1, _, _ = h.values_csc("matern", cp, nug_all, tau)
        x2, _, _ = h.values_csc("matern", cp, _pinned(nug_all.size, nug_all), tau, out=_pinned(nnz))
    want = np.concatenate([r["Lentries"].ravel()[(revNN[:, ::-1] != 0).ravel()], r["Zentries"]])
    assert nf == 0 and np.array_equal(packed, want)
    assert nf2 == 0 and np.array_equal(packed2, want)
    assert np.array_equal(x1, x2)
    # and against the oracle on a sample of rows
    rows = np.r_[0:50, n - 50:n] if layout == "z" else np.r_[n:n + 50, 2 * n - 50:2 * n]
    rc = revCond.astype(np.float64); rc[revCond < 0] = np.nan
    for a, b in ((rows[0], rows[49] + 1), (rows[50], rows[99] + 1)):
        pr = O.RowsProblem(locs2, revNN[a:b], rc[a:b], a, nug_all, "matern", np.array(cp))
        pr.run(2)
        ref = pr.Lentries()
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(r["Lentries"][a:b] - ref) / scale).max() < 1e-9


@pytest.mark.parametrize("locality", ["0", "1"])
def test_chunked_pipeline_with_neighbour_ids_in_any_order(locality, monkeypatch):
    # U_NZentries does not ask for an ordered layout: a row may name ANY location, also later ones.  The staged
    # nugget upload of the chunked call (chunk c brings up what its rows name) must then degrade to "everything
    # before the first chunk", not read nuggets that have not arrived.  Also a row shard, which needs a prefix only.
    monkeypatch.setenv("GPV_LOCALITY", locality)
    n, m = 140000, 6
    rng = np.random.default_rng(11)
    locs = H.make_locs(n, 2, stream=96)
    revNN = H.rev(H.ordered_nn_kdtree(locs, m)).astype(np.int64).copy()       # ordered neighbours, self last ...
    far = rng.integers(0, n, size=n // 50)                                    # ... and in 2 % of the rows a later one
    tgt = np.minimum(far + rng.integers(1, n // 2, size=far.size), n - 1)
    ok = (revNN[far] != (tgt + 1)[:, None]).all(axis=1) & (revNN[far, 0] != 0)
    revNN[far[ok], 0] = tgt[ok] + 1
    revCond = np.zeros_like(revNN, dtype=np.int32)
    nug = H.make_nuggets(n, stream=96)
    cp = [1.0, H.default_range(n, 2), 1.5]
    rc = revCond.astype(np.float64)
    for a, b in ((0, n), (n // 3, n // 3 + 66000)):
        with G.UHandle(locs, revNN, revCond, obs=np.ones(n, dtype=bool), row_begin=a, row_end=b) as h:
            ztail = (a, b) == (0, n)
            assert h.nuggets_read == revNN[a:b].max()             # one past the largest (0-based) id its rows name
            want, nf, _ = h.values_packed("matern", cp, nug, nug, zentries_tail=ztail)
            out = _pinned(want.size)
            out[:] = np.nan
            got, nf2, _ = h.values_packed("matern", cp, nug, nug, zentries_tail=ztail, out=out)
        assert nf == nf2 == 0 and np.array_equal(got, want)
        rows = np.r_[a:a + 40, b - 40:b]
        pr = O.RowsProblem(locs, revNN[rows], rc[rows], 0, nug, "matern", np.array(cp))
        pr.run(2)
        ref = pr.Lentries()
        n0 = (revNN[a:b] != 0).sum(axis=1)
        off = np.concatenate([[0], np.cumsum(n0)])
        for k, r_ in enumerate(rows):
            v = got[off[r_ - a]:off[r_ - a + 1]]
            assert np.abs(v - ref[k, :v.size]).max() <= VAL_TOL * np.abs(ref[k]).max()


@pytest.mark.parametrize("layout", ["z", "zy"])
def test_results_into_pageable_memory_go_through_the_copy_workers(layout):
    # every R vector is pageable: outputs of 8 MB or more are fetched in 4 MB pieces by worker threads of the
    # library (page-locked slots, then memcpy into the caller's buffer).  Same bits as the page-locked route, for
    # the packed vector with its Z tail, the compressed-column values and the N x p matrix of the stateless call.
    n, m = 150000, 10
    locs = H.make_locs(n, 2, stream=95)
    if layout == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        nug_all = np.concatenate([H.make_nuggets(n, stream=95), np.zeros(n)])
    else:
        locs2, NN = locs, H.ordered_nn_kdtree(locs, m)
        Cond, obs = H.layout_yz(NN, "z"), np.ones(n, dtype=bool)
        nug_all = H.make_nuggets(n, stream=95)
    revNN, revCond = H.rev(NN), H.rev(Cond)
    tau = nug_all[:n]
    cp = [1.0, H.default_range(n, 2), 1.5]
    with G.UHandle(locs2, revNN, revCond, obs=obs) as h:
        total = h.packed_len + 2 * n
        assert total * 8 >= (8 << 20)
        want, nf, _ = h.values_packed("matern", cp, nug_all, tau, out=_pinned(total))
        got = np.full(total, np.nan)
        _, nf2, _ = h.values_packed("matern", cp, nug_all, tau, out=got)
        assert nf == nf2 == 0 and np.array_equal(got, want)
        got_nz = np.full(h.packed_len, np.nan)                                # no Z tail
        h.values_packed("matern", cp, nug_all, tau, zentries_tail=False, out=got_nz)
        assert np.array_equal(got_nz, want[:h.packed_len])
        _, nnz, _ = h.csc_dims()
        xw, _, _ = h.values_csc("matern", cp, nug_all, tau, out=_pinned(nnz))
        xg = np.full(nnz, np.nan)
        h.values_csc("matern", cp, nug_all, tau, out=xg)
        assert np.array_equal(xg, xw)
        r = h.U_NZentries("matern", cp, nug_all, tau)                         # row-major -> column-major -> workers
        keep = (revNN[:, ::-1] != 0)
        assert np.array_equal(np.concatenate([r["Lentries"][keep], r["Zentries"]]), want)
        assert not np.any(r["Lentries"][~keep])


def test_stateless_entry_point_recognises_unchanged_inputs_and_changed_ones(monkeypatch):
    # gpv_U_NZentries keeps the handle of its last call (createU calls it hundreds of times with the same locsord /
    # revNNarray) and recognises the arrays by a fingerprint of their CONTENT: an array changed in place, or another
    # revCond alone, must show in the result
    n, m = 3000, 9
    va = _problem(n, m, 2, "SGV", stream=94)
    prep = va["U_prep"]
    nug = H.make_nuggets(n, stream=94)
    cp = [1.1, H.default_range(n, 2), 1.5]
    tol = 1e-8        # SGV blocks condition on the latent field (cond ~ 1e5): this test is about the cache
    call = lambda nnarr, cond: G.U_NZentries(1, n, va["locsord"], nnarr, cond, nug, nug, "matern", cp)
    ref = lambda nnarr, cond: O.U_NZentries(O.max_threads(), n, va["locsord"], nnarr, _rc_double(cond), nug, nug, "matern",
                                            np.asarray(cp, float))
    a, b = call(prep["revNNarray"], prep["revCond"]), call(prep["revNNarray"], prep["revCond"])     # second call: cache hit
    assert np.array_equal(a["Lentries"], b["Lentries"]) and np.array_equal(a["Zentries"], b["Zentries"])
    assert _rowscaled_err(a["Lentries"], ref(prep["revNNarray"], prep["revCond"])["Lentries"]) < tol
    cond2 = prep["revCond"].copy()
    cond2[prep["revNNarray"] != 0] = 0                                       # same neighbours, all conditioned on z
    c = call(prep["revNNarray"], cond2)
    assert not np.array_equal(c["Lentries"], a["Lentries"])
    assert _rowscaled_err(c["Lentries"], ref(prep["revNNarray"], cond2)["Lentries"]) < tol
    nn2 = prep["revNNarray"].copy()
    rows = np.arange(n // 2, n, 7)
    nn2[rows, 0] = 0                                                         # drop the farthest neighbour of some rows
    d = call(nn2, cond2)
    assert _rowscaled_err(d["Lentries"], ref(nn2, cond2)["Lentries"]) < tol
    assert np.all(d["Lentries"][rows, -1] == 0) and np.all(c["Lentries"][rows, -1] != 0)
    e = call(prep["revNNarray"], prep["revCond"])                            # and back
    assert np.array_equal(e["Lentries"], a["Lentries"])
    monkeypatch.setenv("GPV_STATELESS_CACHE", "0")
    f = call(prep["revNNarray"], prep["revCond"])
    assert np.array_equal(f["Lentries"], a["Lentries"])
    G.lib.gpv_release_cached()


def test_whole_loglik_on_gpu_for_pure_z_conditioning():
    # standard Vecchia: denominator terms are per-row closed forms (gpv_loglik_z)
    n, m = 3000, 20
    va = _problem(n, m, 2, "z", stream=98)
    prep = va["U_prep"]
    z = H.make_data(n, stream=98)
    tau = H.make_nuggets(n, stream=98)
    for covType, cp in (("matern", [1.2, 0.06, 1.5]), ("matern", [0.9, 0.05, 0.8]), ("esqe", [0.7, 0.05, 0.4, 0.03])):
        with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
            r = h.loglik_z(covType, cp, tau, tau, z)
        Uo = O.createU(va, cp, tau, covmodel=covType)
        ll_ref = O.vecchia_likelihood_U(z, Uo)
        qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
        assert r["nfail"] == 0
        assert abs(r["quadform_num"] - qr) <= LL_TOL * abs(qr) and abs(r["logdet_num"] - lr) <= LL_TOL * abs(lr)
        assert abs(r["loglik"] - ll_ref) <= LL_TOL * abs(ll_ref), (covType, r["loglik"], ll_ref)
        assert abs(G.vecchia_likelihood(z, va, cp, tau, covmodel=covType) - ll_ref) <= LL_TOL * abs(ll_ref)
    # known answer: m = n-1 => exact Gaussian log-density
    n2 = 31
    locs = H.make_locs(n2, 2, stream=99)
    NN = H.ordered_nn_kdtree(locs, n2 - 1)
    va2 = H.make_vecchia_approx(locs, NN, H.layout_yz(NN, "z"), np.ones(n2, dtype=bool), "z")
    z2 = H.make_data(n2, stream=99)
    ll = G.vecchia_likelihood(z2, va2, [1.3, 0.25, 2.5], 0.2)
    ex = O.exact_loglik(z2, locs, [1.3, 0.25, 2.5], 0.2)
    assert abs(ll - ex) <= LL_TOL * abs(ex)
    # shards: parts add up
    cp = [1.2, 0.06, 1.5]
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
        whole = h.loglik_z("matern", cp, tau, tau, z)
    acc = dict(quadform_num=0.0, logdet_num=0.0, quadform_denom=0.0, logdet_denom=0.0)
    for a, b in ((0, 1100), (1100, 3000)):
        with G.UHandle(va["locsord"], prep["revNNarray"][a:b], prep["revCond"][a:b], obs=va["obs"],
                       row_begin=a, row_end=b) as hs:
            part = hs.loglik_z("matern", cp, tau, tau, z)
            for k in acc:
                acc[k] += part[k]
    for k in acc:
        assert abs(acc[k] - whole[k]) <= 1e-11 * abs(whole[k]), k


def test_loglik_z_refuses_other_layouts():
    va = _problem(400, 6, 2, "SGV", stream=100)
    prep = va["U_prep"]
    tau = np.full(400, 0.1)
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
        with pytest.raises(G.GpvError) as ei:
            h.loglik_z("matern", [1.0, 0.1, 1.5], tau, tau, np.zeros(400))
        assert ei.value.status == 5


def test_cfg2_full_size_properties():
    # BASELINE configs[1] at full size (n = 1e6, m = 30): the oracle cannot run 1e6 rows in seconds, so
    # check a row sample against it and size-independent properties on everything
    n, m = 1_000_000, 30
    locs = H.make_locs(n, 2, stream=2)
    revNN = H.ordered_nn_gpu(locs, m)
    revCond = np.zeros(revNN.shape, dtype=np.int32)
    revCond[revNN == 0] = np.iinfo(np.int32).min
    revCond[:, -1] = 1
    nug = H.make_nuggets(n, stream=2)
    z = H.make_data(n, stream=2)
    rng_ = H.default_range(n, 2)
    with G.UHandle(locs, revNN, revCond, obs=np.ones(n, dtype=np.int32)) as h:
        a = h.U_NZentries("matern", [1.0, rng_, 1.5], nug, nug)
        a2 = h.U_NZentries("matern", [1.0, rng_, 1.5], nug, nug)
        packed, nf, _ = h.values_packed("matern", [1.0, rng_, 1.5], nug, nug, zentries_tail=False)
        b = h.U_NZentries("matern", [4.0, rng_, 1.5], 4.0 * nug, 4.0 * nug)
        ll = h.loglik_z("matern", [1.0, rng_, 1.5], nug, nug, z)
    L = a["Lentries"]
    assert a["nfail"] == 0 and nf == 0
    assert np.array_equal(L, a2["Lentries"])                                   # idempotent, deterministic
    assert np.array_equal(packed, L.ravel()[(revNN[:, ::-1] != 0).ravel()])    # packed a9 order
    n0 = (revNN != 0).sum(axis=1)
    diag = L[np.arange(n), n0 - 1]
    assert np.all(diag > 0) and np.all(np.isfinite(L))
    assert np.all(L[n0 < m + 1][:, -1] == 0)                                   # zero fill beyond n0
    assert _rowscaled_err(b["Lentries"] * 2.0, L) < 1e-12                      # cov -> 4 cov => U -> U/2
    # logdet.num from the U values vs the fused reduction
    ld = -2.0 * np.log(diag).sum() + np.log(nug).sum()
    assert abs(ld - ll["logdet_num"]) <= 1e-10 * abs(ld)
    # row sample against the oracle
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(n, 3000, replace=False))
    rc = revCond[rows].astype(np.float64)
    rc[revCond[rows] < 0] = np.nan
    pr = O.RowsProblem(locs, revNN[rows], rc, 0, nug, "matern", np.array([1.0, rng_, 1.5]))
    pr.run(O.max_threads())
    assert np.array_equal(pr.Lentries() == 0, L[rows] == 0)
    assert _rowscaled_err(L[rows], pr.Lentries()) < VAL_TOL


@pytest.mark.parametrize("n,m,d,covType,cp_tail", [
    (4_000_000, 30, 2, "matern", [0.8]),            # BASELINE configs[2] shape: general-nu branch, locality layer on
    (1_000_000, 40, 3, "esqe", [0.5, None]),        # BASELINE configs[3] shape: 16-lane groups, 3-D, esqe
])
def test_cfg3_cfg4_large_size_properties(n, m, d, covType, cp_tail):
    """The other large BASELINE shapes at sizes the oracle cannot run in seconds: size-independent properties on
    everything (idempotence, packed order, the scale property, fused sums against the values, zero fill) and a
    3 000-row sample against the oracle.  n = 4e6 keeps the GPU suite short; bench.py runs cfg3 at its full 1e7 with
    the same sample check inside the run."""
    locs = H.make_locs(n, d, stream=3)
    revNN = H.ordered_nn_gpu(locs, m)
    revCond = np.zeros(revNN.shape, dtype=np.int32)
    revCond[revNN == 0] = np.iinfo(np.int32).min
    revCond[:, -1] = 1
    nug = H.make_nuggets(n, stream=3)
    z = H.make_data(n, stream=3)
    rng_ = H.default_range(n, d)
    cp = [1.0, rng_] + [rng_ if v is None else v for v in cp_tail]
    cp4 = list(cp)
    cp4[0] *= 4.0
    if covType == "esqe":
        cp4[2] *= 4.0
    with G.UHandle(locs, revNN, revCond, obs=np.ones(n, dtype=np.int32)) as h:
        a = h.U_NZentries(covType, cp, nug, nug)
        name = h.last_kernel_name()
        packed, nf, _ = h.values_packed(covType, cp, nug, nug, zentries_tail=False)
        b = h.U_NZentries(covType, cp4, 4.0 * nug, 4.0 * nug)
        ll = h.loglik_z(covType, cp, nug, nug, z)
    assert name.startswith("u_band<G=%d" % (8 if m == 30 else 16)) and ("general" in name) == (covType == "matern")
    L = a["Lentries"]
    assert a["nfail"] == 0 and nf == 0 and ll["nfail"] == 0
    assert np.array_equal(packed, L.ravel()[(revNN[:, ::-1] != 0).ravel()])    # packed a9 order (chunked pipeline, per-chunk order)
    n0 = (revNN != 0).sum(axis=1)
    diag = L[np.arange(n), n0 - 1]
    assert np.all(diag > 0) and np.all(np.isfinite(L))
    assert np.all(L[n0 < m + 1][:, -1] == 0)
    assert _rowscaled_err(b["Lentries"] * 2.0, L) < 1e-12                      # cov -> 4 cov => U -> U/2
    ld = -2.0 * np.log(diag).sum() + np.log(nug).sum()
    assert abs(ld - ll["logdet_num"]) <= 1e-10 * abs(ld)
    rng = np.random.default_rng(1)
    rows = np.sort(rng.choice(np.arange(m + 1, n), 3000, replace=False))
    rc = revCond[rows].astype(np.float64)
    rc[revCond[rows] < 0] = np.nan
    pr = O.RowsProblem(locs, revNN[rows], rc, 0, nug, covType, np.array(cp))
    pr.run(O.max_threads())
    assert _rowscaled_err(L[rows], pr.Lentries()) < VAL_TOL


def test_randomised_sweep_against_oracle():
    # seeded random shapes: set size, dimension, covariance, conditioning mask, missing pattern,
    # nugget vector -- every case against the CPU restatement (values) and its failure count
    rng = np.random.default_rng(20240601)
    checked = 0
    for case in range(24):
        m = int(rng.integers(1, 64))
        d = int(rng.integers(1, 6))
        n = int(rng.integers(m + 2, 400))
        locs = rng.random((n, d))
        NN = O.find_ordered_nn_brute(locs, m) if n < 250 else H.ordered_nn_kdtree(locs, m)
        Cond = -np.ones(NN.shape, dtype=np.int8)
        nz = NN != 0
        Cond[nz] = (rng.random(int(nz.sum())) < 0.5).astype(np.int8)       # arbitrary y/z mix
        Cond[:, 0] = 1
        revNN, revCond = H.rev(NN).copy(), H.rev(Cond).copy()
        if case % 3 == 0:                                                   # holes anywhere (find() compaction)
            kill = rng.random(revNN.shape) < 0.1
            kill[:, -1] = False
            revNN[kill] = 0
        nug = 0.02 + 0.3 * rng.random(n)
        kind = case % 5
        rng_ = 0.05 + 0.4 * rng.random()
        if kind == 4:
            covType, cp = "esqe", [0.3 + rng.random(), rng_, 0.2 + rng.random(), 0.5 * rng_]
        else:
            covType, cp = "matern", [0.5 + rng.random(), rng_, [0.5, 1.5, 2.5, float(0.2 + 3 * rng.random())][kind]]
        rc = revCond.astype(np.float64)
        rc[revCond < 0] = np.nan
        got = G.U_NZentries(1, n, locs, revNN, revCond, nug, nug, covType, cp)
        ref = O.U_NZentries(2, n, locs, revNN, rc, nug, nug, covType, np.array(cp), mode=2)   # __float128 arbiter
        ref64 = O.U_NZentries(2, n, locs, revNN, rc, nug, nug, covType, np.array(cp), mode=1)
        # a block can be numerically singular in fp64 (1-D, smooth kernel, latent neighbours without
        # nugget): whether a pivot then rounds to <= 0 is implementation dependent, so a failed row is
        # accepted only if the oracle's conditioning proxy says the block is singular to fp64
        n0 = (revNN != 0).sum(axis=1)
        gfail = np.nonzero((got["Lentries"] == 0).all(axis=1) & (n0 > 0))[0]
        assert gfail.size == got["nfail"]
        for k in gfail:
            assert O.block_cond_proxy(int(k), locs, revNN, rc, nug, covType, np.array(cp)) > 1e12, (case, int(k))
        # values are compared on the rows the fp64 restatement itself resolves to 1e-12 against quad
        # precision (condition number up to ~1e4); on worse-conditioned blocks two correct fp64
        # evaluations of the covariance (1e-15 apart) already differ by cond * 1e-15 in U
        scale = np.abs(ref["Lentries"]).max(axis=1)
        scale[scale == 0] = 1.0
        row_ref = np.abs(ref64["Lentries"] - ref["Lentries"]).max(axis=1) / scale
        row_gpu = np.abs(got["Lentries"] - ref["Lentries"]).max(axis=1) / scale
        ok = row_ref < 1e-12
        ok[gfail] = False
        assert ok.sum() >= 3, case
        assert np.array_equal(got["Lentries"][ok] == 0, ref["Lentries"][ok] == 0), case
        assert row_gpu[ok].max() < VAL_TOL, (case, m, d, covType, cp, float(row_gpu[ok].max()))
        checked += int(ok.sum())
    assert checked > 2000


@pytest.mark.parametrize("layout", ["z", "zy"])
def test_single_process_multi_device_front_end(layout):
    # gpv_multi_*: worker threads, one shard per entry of `devices` (an ordinal may repeat, so the
    # slicing logic is exercised on a one-GPU box; on an N-GPU box pass range(N))
    n, m = 5000, 12
    locs = H.make_locs(n, 2, stream=101)
    if layout == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        nug_all = np.concatenate([H.make_nuggets(n, stream=101), np.zeros(n)])
    else:
        locs2, NN = locs, H.ordered_nn_kdtree(locs, m)
        Cond, obs = H.layout_yz(NN, "z"), np.ones(n, dtype=bool)
        nug_all = H.make_nuggets(n, stream=101)
    revNN, revCond = H.rev(NN), H.rev(Cond)
    tau, z = nug_all[:n], H.make_data(n, stream=101)
    cp = [1.0, H.default_range(n, 2), 1.5]
    skip = n if layout == "zy" else 0
    with G.UHandle(locs2, revNN, revCond, obs=obs) as h:
        want, nf, _ = h.values_packed("matern", cp, nug_all, tau)
        q, l, _ = h.loglik_numerator("matern", cp, nug_all, tau, z, skip_rows=skip)
        llz = h.loglik_z("matern", cp, nug_all, tau, z) if layout == "z" else None
        want_p, want_i = h.csc_pattern()
        want_x, _, _ = h.values_csc("matern", cp, nug_all, tau)
    ndev = G.lib.gpv_device_count()
    # the oracle's answer in the packed (createU.R:158-160) order: every device list below -- all GPUs of the box
    # among them -- is held to the ORACLE, not only to the one-GPU output
    ref = O.U_NZentries(O.max_threads(), n, locs2, revNN, _rc_double(revCond), nug_all, tau, "matern", np.array(cp))
    keep = (revNN[:, ::-1] != 0).ravel()       # Lentries holds a row's n0 values in its FIRST n0 slots
    ref_packed = np.concatenate([ref["Lentries"].ravel()[keep], ref["Zentries"]])
    ref_scale = np.concatenate([np.repeat(np.abs(ref["Lentries"]).max(axis=1), revNN.shape[1])[keep], np.abs(ref["Zentries"])])
    if layout == "zy":
        arb = O.U_NZentries(O.max_threads(), n, locs2, revNN, _rc_double(revCond), nug_all, tau, "matern", np.array(cp), mode=2)
        arb_packed = np.concatenate([arb["Lentries"].ravel()[keep], arb["Zentries"]])
        err_ref = float((np.abs(ref_packed - arb_packed) / ref_scale).max())
    for devices in ([0], [0, 0, 0], list(range(ndev)), list(range(ndev)) * 2):
        with G.MultiHandle(locs2, revNN, revCond, obs=obs, devices=devices) as mh:
            assert mh.row_cuts[0] == 0 and mh.row_cuts[-1] == locs2.shape[0] and mh.packed_len == want.size - 2 * n
            got, nf2, _ = mh.values_packed("matern", cp, nug_all, tau)
            assert nf2 == nf and np.array_equal(got, want)
            if layout == "zy":       # ill-conditioned zy blocks: as close to the __float128 arbiter as the fp64 oracle is
                assert float((np.abs(got - arb_packed) / ref_scale).max()) < max(VAL_TOL, 3 * err_ref)
            else:
                assert float((np.abs(got - ref_packed) / ref_scale).max()) < VAL_TOL
            got_p, got_i = mh.csc_pattern()
            got_x, nf3, _ = mh.values_csc("matern", cp, nug_all, tau)
            assert nf3 == nf and np.array_equal(got_p, want_p) and np.array_equal(got_i, want_i) and np.array_equal(got_x, want_x)
            q2, l2, _ = mh.loglik_numerator("matern", cp, nug_all, tau, z, skip_rows=skip)
            assert abs(q2 - q) <= 1e-12 * abs(q) and abs(l2 - l) <= 1e-12 * abs(l)
            if layout == "z":
                r = mh.loglik_z("matern", cp, nug_all, tau, z)
                assert abs(r["loglik"] - llz["loglik"]) <= 1e-12 * abs(llz["loglik"])
            if layout == "zy" and len(devices) == 3:
                # the n dummy rows are free: the first cut lies beyond them
                assert mh.row_cuts[1] > n


@pytest.mark.parametrize("layout", ["z", "zy", "SGV"])
def test_locality_layer_changes_the_order_of_work_not_the_results(layout, monkeypatch):
    """Large handles keep a Morton-sorted replica of the per-location data and process the rows of every output
    chunk in Morton order (gpv_capi.cu, locality layer; automatic above ~2e6 locations).  GPV_LOCALITY=1 forces it
    on a small problem: U values (row-major, packed, chunked pipeline, compressed-column), failure reporting and
    the likelihood sums must equal the plain path -- values bit for bit, sums to reduction-order rounding."""
    n, m = 70000, 12                    # >= 65536 sets: the chunked packed pipeline is active
    locs = H.make_locs(n, 2, stream=131)
    z = H.make_data(n, stream=131)
    if layout == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        nug_all = np.concatenate([H.make_nuggets(n, stream=131), np.zeros(n)])
        skip = n
    else:
        locs2, NN = locs, H.ordered_nn_kdtree(locs, m)
        Cond = H.layout_yz(NN, "z") if layout == "z" else H.whichCondOnLatent(NN)
        obs = np.ones(n, dtype=bool)
        nug_all = H.make_nuggets(n, stream=131)
        skip = 0
    nug_all[1234] = -30.0               # a few failing rows: nfail / first_fail must agree as well
    revNN, revCond = H.rev(NN), H.rev(Cond)
    tau = np.abs(nug_all[:n])
    res = {}
    for loc in ("0", "1"):
        monkeypatch.setenv("GPV_LOCALITY", loc)
        out = {}
        for covType, cp in (("matern", [1.0, H.default_range(n, 2), 1.5]), ("matern", [1.0, H.default_range(n, 2), 0.8])):
            with G.UHandle(locs2, revNN, revCond, obs=obs) as h:
                r = h.U_NZentries(covType, cp, nug_all, tau)
                packed, nf, ff = h.values_packed(covType, cp, nug_all, tau)
                ll = h.loglik_numerator(covType, cp, nug_all, tau, z, skip_rows=skip)
                try:
                    csc = h.values_csc(covType, cp, nug_all, tau)[0]
                except G.GpvError:          # layouts whose sets hold a U row twice go the triplet route
                    csc = None
            out[cp[2]] = (r["Lentries"], r["nfail"], r["first_fail"], packed, nf, ff, ll, csc)
        res[loc] = out
    for nu in res["0"]:
        a, b = res["0"][nu], res["1"][nu]
        assert a[1] == b[1] > 0 and a[2] == b[2] and a[4] == b[4] and a[5] == b[5]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
        if a[7] is not None:
            assert b[7] is not None and np.array_equal(a[7], b[7])
        assert np.allclose(np.asarray(a[6][:2], dtype=float), np.asarray(b[6][:2], dtype=float), rtol=1e-12, atol=0)
        assert a[6][2] == b[6][2]


def test_distributed_likelihood_entry_point_on_one_rank():
    """gpv_loglik_z_dist (gpv_dist.inc) with a one-rank NCCL communicator: the slices are the whole vectors, the
    broadcasts and the all-reduce run on one device -- every line of the multi-process path except a second GPU
    (tools/dist_check.py runs it under torchrun on 2+ GPUs).  Must equal gpv_loglik_z."""
    n, m = 20000, 15
    va = _problem(n, m, 2, "z", stream=140)
    prep = va["U_prep"]
    z = H.make_data(n, stream=140)
    tau = H.make_nuggets(n, stream=140)
    cp = [1.1, H.default_range(n, 2), 0.8]
    with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=va["obs"]) as h:
        want = h.loglik_z("matern", cp, tau, tau, z)
        try:
            h.dist_init(G.UHandle.dist_unique_id(), 0, 1)
        except G.GpvError as e:
            if e.status == 5:
                pytest.skip("NCCL not loadable on this box")
            raise
        cuts = np.array([0, n], dtype=np.int64)
        got = h.loglik_z_dist("matern", cp, tau, tau, z, cuts, cuts)
        again = h.loglik_z_dist("matern", [1.3, cp[1], 1.5], None, None, None, None, None)   # resident data, new covparms
        want2 = h.loglik_z("matern", [1.3, cp[1], 1.5], tau, tau, z)
    for k in ("loglik", "quadform_num", "logdet_num", "quadform_denom", "logdet_denom"):
        assert abs(got[k] - want[k]) <= 1e-13 * abs(want[k]), k
        assert abs(again[k] - want2[k]) <= 1e-13 * abs(want2[k]), k
    assert got["nfail"] == want["nfail"] == 0
