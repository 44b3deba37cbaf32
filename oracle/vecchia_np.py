"""numpy restatement of the R side of the path.  TEST INFRASTRUCTURE ONLY.

Follows, loop for loop, R/U_sparsity.R:5-81, R/createU.R:65-201 (non-MRA branch),
R/vecchia_likelihood.R:14-99, R/vecchia_prediction.R:62-111 (U2V),
R/vecchia_specify.R:29-240 (the 'NN' conditioning subset) and R/whichCondOnLatent.R:2-27.
Pure-Python loops: small n only (tests use n <= a few thousand).

Conventions (R objects -> numpy):
  * location ids are 1-based exactly as in R; NA in NNarray is stored as 0
    (createU.R:146-147 does the same replacement before the .Call);
  * Cond / revCond : int8, 1 = TRUE (condition on latent y), 0 = FALSE (on response z), -1 = NA;
  * ord, ord.z : 1-based permutations;
  * sparse U : scipy.sparse.csc_matrix (R: Matrix::dgCMatrix); duplicates are summed by both.
"""
import numpy as np
import scipy.sparse as sp

from . import ref_c

NA = 0


# ----------------------------------------------------------------------------------------------
# R/U_sparsity.R:5-81
# ----------------------------------------------------------------------------------------------
def U_sparsity(locs, NNarray, obs, Cond):
    NNarray = np.asarray(NNarray, dtype=np.int64)
    Cond = np.asarray(Cond, dtype=np.int8)
    obs = np.asarray(obs, dtype=bool)
    nnp = locs.shape[0]                      # :10
    n = int(obs.sum())                       # :11
    size = nnp + n                           # :12
    nentries = int((NNarray != NA).sum())    # :15

    cur = 1                                  # :19-29
    latent_map = np.zeros(nnp, dtype=np.int64)
    observed_map = np.zeros(nnp, dtype=np.int64)   # 0 stands for R's NA
    for k in range(nnp):
        latent_map[k] = cur
        cur += 1
        if obs[k]:
            observed_map[k] = cur
            cur += 1

    revNNarray = NNarray[:, ::-1].copy()     # :32
    revCondOnLatent = Cond[:, ::-1].copy()   # :33

    rowpointers = np.zeros(nentries, dtype=np.int64)   # :36-56
    colindices = np.zeros(nentries, dtype=np.int64)
    rowpointers[0] = colindices[0] = 1
    cur = 0
    for k in range(nnp):
        inds = revNNarray[k, :]
        keep = inds != NA
        inds0 = inds[keep]
        n0 = inds0.size
        revCond = revCondOnLatent[k, keep] == 1
        cur_row = latent_map[k]
        cur_cols = np.zeros(n0, dtype=np.int64)
        cur_cols[revCond] = latent_map[inds0[revCond] - 1]
        cur_cols[~revCond] = observed_map[inds0[~revCond] - 1]
        if k > 0:
            rowpointers[cur:cur + n0] = cur_row
            colindices[cur:cur + n0] = cur_cols
        cur += n0

    Zrowpointers = np.zeros(2 * n, dtype=np.int64)      # :59-69
    Zcolindices = np.zeros(2 * n, dtype=np.int64)
    cur = 0
    for k in range(nnp):
        if obs[k]:
            Zrowpointers[cur:cur + 2] = observed_map[k]
            Zcolindices[cur:cur + 2] = (latent_map[k], observed_map[k])
            cur += 2

    return dict(revNNarray=revNNarray, revCond=revCondOnLatent, n_cores=ref_c.max_threads(),
                size=size,
                rowpointers=np.concatenate([rowpointers, Zrowpointers]),
                colindices=np.concatenate([colindices, Zcolindices]),
                y_ind=latent_map, observed_map=observed_map)


# ----------------------------------------------------------------------------------------------
# R/whichCondOnLatent.R:2-27  (SGV rule)
# ----------------------------------------------------------------------------------------------
def whichCondOnLatent(NNarray, firstind_pred=None):
    NNarray = np.asarray(NNarray, dtype=np.int64)
    n, p = NNarray.shape
    m = p - 1
    if firstind_pred is None:
        firstind_pred = n + 1
    C = -np.ones((n, p), dtype=np.int8)
    C[0, 0] = 1

    def prod_row(l):  # NNarray[l,]*CondOnLatent[l,] with NA propagated; returns the set of values
        r = NNarray[l - 1]
        c = C[l - 1]
        vals = []
        for a, b in zip(r, c):
            if a == NA or b < 0:
                continue       # NA never matches a non-NA element in is.element
            vals.append(int(a) * int(b))
        return set(vals)

    for k in range(2, n + 1):
        row = NNarray[k - 1]
        latents = np.zeros(p, dtype=np.int64)        # rep(0,m) extended by assignment to index m+1
        for ind in range(2, m + 2):
            l = row[ind - 1]
            if l != NA and l < firstind_pred:
                s = prod_row(l)
                latents[ind - 1] = sum(1 for a in row if (a != NA and int(a) in s))
        ind = row[int(np.argmax(latents == latents.max()))]
        if ind == NA:
            # cannot happen for NN layouts (first max is at a valid neighbour or at self)
            s = set()
        else:
            s = prod_row(ind)
        for j in range(p):
            a = row[j]
            if a == NA:
                C[k - 1, j] = -1
            else:
                C[k - 1, j] = 1 if int(a) in s else 0
        for j in range(p):
            if row[j] != NA and row[j] >= firstind_pred:
                C[k - 1, j] = 1
        C[k - 1, 0] = 1
        for j in range(p):
            if row[j] == NA:
                C[k - 1, j] = -1
    return C


def U_NZentries_mat(n, revNNarray, covVals, nuggets_obsord):
    """src/U_NZentries.cpp:126-197 restated (numpy, small cases): covmat = covVals(inds, inds), no nugget,
    revCond unused; chol(., "upper") + solve(R, e_last); a failing Cholesky leaves the row zero."""
    rnn = np.asarray(revNNarray)
    N, p = rnn.shape
    L = np.zeros((N, p))
    nfail = 0
    for k in range(N):
        inds = rnn[k][(rnn[k] != 0) & (rnn[k] != NA)] - 1
        n0 = inds.size
        if n0 == 0:
            continue
        cm = np.asarray(covVals)[np.ix_(inds, inds)]
        try:
            if not np.all(np.isfinite(cm)):
                raise np.linalg.LinAlgError
            R = np.linalg.cholesky(cm).T
        except np.linalg.LinAlgError:
            nfail += 1
            continue
        e = np.zeros(n0); e[-1] = 1.0
        L[k, :n0] = np.linalg.solve(R, e)
    tau = np.asarray(nuggets_obsord, dtype=np.float64)
    Z = np.empty(2 * n)
    Z[0::2] = -1.0 / np.sqrt(tau)
    Z[1::2] = 1.0 / np.sqrt(tau)
    return dict(Lentries=L, Zentries=Z, nfail=nfail)


# ----------------------------------------------------------------------------------------------
# neighbour search stand-ins (GpGp::find_ordered_nn, FNN::get.knn are third-party inputs)
# ----------------------------------------------------------------------------------------------
def find_ordered_nn_brute(locsord, m):
    """Row i (1-based) = (i, its min(m, i-1) nearest among rows < i, nearest first), NA-padded.
    Semantics of GpGp::find_ordered_nn as used at R/vecchia_specify.R:159."""
    locsord = np.asarray(locsord, dtype=np.float64)
    N = locsord.shape[0]
    NN = np.zeros((N, m + 1), dtype=np.int64)
    for i in range(N):
        NN[i, 0] = i + 1
        if i == 0:
            continue
        dd = np.sqrt(((locsord[:i] - locsord[i]) ** 2).sum(axis=1))
        o = np.argsort(dd, kind="stable")[:m]
        NN[i, 1:1 + o.size] = o + 1
    return NN


def _get_knn_brute(x, k):
    """FNN::get.knn(x, k)$nn.index : k nearest *other* points, nearest first (1-based)."""
    n = x.shape[0]
    out = np.zeros((n, k), dtype=np.int64)
    for i in range(n):
        dd = np.sqrt(((x - x[i]) ** 2).sum(axis=1))
        dd[i] = np.inf
        out[i] = np.argsort(dd, kind="stable")[:k] + 1
    return out


# ----------------------------------------------------------------------------------------------
# R/vecchia_specify.R:29-240, conditioning='NN' subset.  ordering: 'none' or an explicit
# 1-based permutation (MaxMin and coord orderings are input producers outside the path).
# ----------------------------------------------------------------------------------------------
def vecchia_specify(locs, m, ordering="none", cond_yz=None, locs_pred=None, ord_pred=None):
    locs = np.asarray(locs, dtype=np.float64)
    n, spatial_dim = locs.shape
    if m > n:
        m = n - 1
    have_pred = locs_pred is not None
    if cond_yz is None:                                   # :92-96
        cond_yz = "SGV" if (not have_pred or spatial_dim == 1) else "zy"

    if m == 0:                                            # :59-72
        ord_ = np.arange(1, n + 1)
        NNarray = np.stack([ord_, np.zeros(n, dtype=np.int64)], axis=1)
        Cond = -np.ones((n, 2), dtype=np.int8)
        Cond[:, 0] = 1
        obs = np.ones(n, dtype=bool)
        U_prep = U_sparsity(locs, NNarray, obs, Cond)
        return dict(locsord=locs.copy(), obs=obs, ord=ord_, ord_z=ord_, ord_pred="general",
                    U_prep=U_prep, cond_yz="false", conditioning="NN", ic0=False)

    if isinstance(ordering, str):
        if ordering != "none":
            raise NotImplementedError("oracle supports ordering='none' or an explicit permutation")
        ord_obs = np.arange(1, n + 1)
    else:
        ord_obs = np.asarray(ordering, dtype=np.int64)

    if not have_pred:                                     # :100-118
        ord_ = ord_obs
        ord_z = ord_
        locsord = locs[ord_ - 1]
        obs = np.ones(n, dtype=bool)
        ordering_pred = "general"
        n_p = 0
    else:                                                 # :120-149 (obspred ordering)
        locs_pred = np.asarray(locs_pred, dtype=np.float64)
        n_p = locs_pred.shape[0]
        locs_all = np.vstack([locs, locs_pred])
        if ord_pred is None:
            ord_pred = np.arange(1, n_p + 1)
        ord_ = np.concatenate([ord_obs, np.asarray(ord_pred, dtype=np.int64) + n])
        ord_z = ord_obs
        locsord = locs_all[ord_ - 1]
        obs = np.concatenate([np.ones(n, dtype=bool), np.zeros(n_p, dtype=bool)])[ord_ - 1]
        ordering_pred = "obspred"

    NNarray = find_ordered_nn_brute(locsord, m)           # :159

    if cond_yz == "SGV":                                  # :182-183
        Cond = whichCondOnLatent(NNarray, firstind_pred=n + 1)
    elif cond_yz == "y":                                  # :186-188
        Cond = -np.ones(NNarray.shape, dtype=np.int8)
        Cond[NNarray != NA] = 1
    elif cond_yz == "z":                                  # :189-190
        Cond = -np.ones(NNarray.shape, dtype=np.int8)
        Cond[NNarray != NA] = 0
        Cond[:, 0] = 1
    elif cond_yz in ("RVP", "LK", "zy"):                  # :191-224
        obs = np.concatenate([np.ones(n, dtype=bool), np.zeros(locsord.shape[0], dtype=bool)])
        locsord = np.vstack([locsord[:n], locsord])
        NNs = _get_knn_brute(locsord[:n], m - 1)
        if cond_yz in ("RVP", "zy"):
            prev = NNs < np.arange(1, n + 1)[:, None]
            NNs[prev] += n
        NNarray_z = np.zeros((n, m + 1), dtype=np.int64)
        NNarray_z[:, 0] = np.arange(1, n + 1)
        NNarray_y = np.concatenate([np.arange(1, n + 1)[:, None] + n,
                                    np.arange(1, n + 1)[:, None], NNs], axis=1)
        if not have_pred:
            NNarray_yp = np.zeros((0, m + 1), dtype=np.int64)
            ordering_pred = "obspred"
        else:
            NNarray_yp = NNarray[n:n + n_p].copy()
            if cond_yz == "zy":
                NNarray_yp[NNarray_yp != NA] += n
            else:
                NNarray_yp[NNarray_yp > n] += n
        NNarray = np.vstack([NNarray_z, NNarray_y, NNarray_yp])
        Cond = -np.ones(NNarray.shape, dtype=np.int8)
        Cond[NNarray != NA] = (NNarray[NNarray != NA] > n).astype(np.int8)
        Cond[:, 0] = 1
        cond_yz = "zy"
    else:
        raise ValueError(f"cond.yz='{cond_yz}' not defined")

    U_prep = U_sparsity(locsord, NNarray, obs, Cond)      # :230
    return dict(locsord=locsord, obs=obs, ord=ord_, ord_z=ord_z, ord_pred=ordering_pred,
                U_prep=U_prep, cond_yz=cond_yz, ic0=False, conditioning="NN", NNarray=NNarray,
                Cond=Cond)


# ----------------------------------------------------------------------------------------------
# R/createU.R:65-201 (non-MRA branch :141-163)
# ----------------------------------------------------------------------------------------------
def createU(vecchia_approx, covparms, nuggets, covmodel="matern", U_NZentries=None, mode=0):
    """U_NZentries: callable with the reference's nine arguments; defaults to the C++ oracle."""
    va = vecchia_approx
    prep = va["U_prep"]
    obs = np.asarray(va["obs"], dtype=bool)
    n = int(obs.sum())                                            # :67
    size = prep["size"]                                           # :68
    latent = np.isin(np.arange(1, size + 1), prep["y_ind"])       # :69
    ord_ = np.asarray(va["ord"])
    nuggets = np.atleast_1d(np.asarray(nuggets, dtype=np.float64))
    if nuggets.size == 1:                                         # :74
        nuggets = np.repeat(nuggets, n)
    nuggets_all = np.concatenate([nuggets, np.zeros(int(latent.sum()) - n)])   # :75
    if va["cond_yz"] == "zy":                                     # :76
        ord_all = np.concatenate([ord_[:n], ord_ + n])
    else:
        ord_all = ord_
    nuggets_all_ord = nuggets_all[ord_all - 1]                    # :77
    nuggets_ord = nuggets_all[np.asarray(va["ord_z"]) - 1]        # :78
    zero_nuggets = bool(np.any(nuggets == 0))                     # :79

    revNN = prep["revNNarray"].copy()
    revCond = prep["revCond"].copy()
    if zero_nuggets:                                              # :83-86
        zero_ids = np.nonzero(nuggets_ord == 0)[0] + 1
        zero_cond = np.isin(revNN, zero_ids) & (revNN != NA)
        revCond[zero_cond] = 1

    if not isinstance(covmodel, str):
        raise NotImplementedError("oracle covers the character covmodel branch (createU.R:152)")
    rc_double = revCond.astype(np.float64)
    rc_double[revCond < 0] = np.nan
    if U_NZentries is None:
        def U_NZentries(*a):
            return ref_c.U_NZentries(*a, mode=mode)
    U_entries = U_NZentries(prep["n_cores"], n, va["locsord"], revNN, rc_double,      # :152-154
                            nuggets_all_ord, nuggets_ord, covmodel, np.asarray(covparms, float))

    # :158-160  (apply(revNNarray,1,rev) is the un-reversed NNarray; its non-NA slots select the
    # first n0 entries of every Lentries row)
    not_na = (prep["revNNarray"][:, ::-1] != NA).ravel()
    Lentries = np.asarray(U_entries["Lentries"]).ravel()[not_na]
    allLentries = np.concatenate([Lentries, np.asarray(U_entries["Zentries"]).ravel()])
    U = sp.coo_matrix((allLentries, (prep["colindices"] - 1, prep["rowpointers"] - 1)),   # :161-162
                      shape=(size, size)).tocsc()
    U.sum_duplicates()

    if va["cond_yz"] == "zy":                                     # :166-171
        dummy = 2 * np.arange(1, n + 1) - 1
        keep = np.ones(size, dtype=bool)
        keep[dummy - 1] = False
        U = U[keep][:, keep]
        latent = latent[keep]
        keep_obs = np.ones(obs.size, dtype=bool)
        keep_obs[n:2 * n] = False
        obs = obs[keep_obs]

    zero_nugg = {}
    if zero_nuggets:                                              # :174-193
        if va["cond_yz"] == "zy":
            raise NotImplementedError("zy + zero nuggets relies on R recycling semantics")
        diagU = U.diagonal()
        inds_U = np.nonzero(np.isinf(diagU) & (diagU > 0))[0] + 1
        Ud = U.toarray()
        cond_on = np.array([np.nonzero(Ud[:, j - 1] != 0)[0].min() + 1 for j in inds_U])
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
                ic0=va.get("ic0", False), U_entries=U_entries, nuggets_all_ord=nuggets_all_ord,
                nuggets_ord=nuggets_ord)


def _revMat(a):
    return a[::-1, ::-1]


# ----------------------------------------------------------------------------------------------
# R/vecchia_prediction.R:62-111, dense algebra (small n)
# ----------------------------------------------------------------------------------------------
def U2V(U_obj):
    U = U_obj["U"].toarray()
    latent = U_obj["latent"]
    U_y = U[latent, :]
    if U_obj["cond_yz"] == "zy":                              # :68-70
        return _revMat(U_y[:, latent])
    if U_obj["ord_pred"] != "obspred":                        # :72-83
        W = U_y @ U_y.T
        return np.linalg.cholesky(_revMat(W))                 # t(chol(W.rev)) = lower factor
    last_obs = int(np.nonzero(~latent)[0].max()) + 1          # :86-108
    latents_before = int(latent[:last_obs].sum())
    latents_after = int(latent[last_obs:].sum())
    V_pr = _revMat(U_y[:, last_obs:])
    U_oo = U_y[:latents_before, :last_obs]
    A = U_oo @ U_oo.T
    V_oor = np.linalg.cholesky(_revMat(A))
    V_or = np.vstack([np.zeros((latents_after, latents_before)), V_oor])
    return np.hstack([V_pr, V_or])


# ----------------------------------------------------------------------------------------------
# R/vecchia_likelihood.R:63-99
# ----------------------------------------------------------------------------------------------
def loglik_numerator_from_U(z, U_obj):
    """(quadform.num, logdet.num) of R/vecchia_likelihood.R:74-76, from an assembled U."""
    U = U_obj["U"].tocsr()
    latent = U_obj["latent"]
    zord = np.asarray(z, dtype=np.float64)[U_obj["ord_z"] - 1]
    z1 = U[np.nonzero(~latent)[0], :].T @ zord
    quadform_num = float(np.sum(z1 ** 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        logdet_num = float(-2 * np.sum(np.log(U.diagonal())))
    return quadform_num, logdet_num, z1


def vecchia_likelihood_U(z, U_obj):
    latent = U_obj["latent"]
    const = float((~latent).sum() * np.log(2 * np.pi))        # :71
    quadform_num, logdet_num, z1 = loglik_numerator_from_U(z, U_obj)   # :74-76
    if latent.sum() == 0:                                     # :79-81
        logdet_denom = quadform_denom = 0.0
    else:                                                     # :83-91
        U_y = U_obj["U"].tocsr()[np.nonzero(latent)[0], :]
        z2 = np.asarray(U_y @ z1).ravel()
        V_ord = U2V(U_obj)
        z3 = np.linalg.solve(V_ord, z2[::-1])
        quadform_denom = float(np.sum(z3 ** 2))
        logdet_denom = float(-2 * np.sum(np.log(np.diag(V_ord))))
    neg2loglik = logdet_num - logdet_denom + quadform_num - quadform_denom + const   # :95
    return -neg2loglik / 2


def vecchia_likelihood(z, vecchia_approx, covparms, nuggets, covmodel="matern", **kw):
    """R/vecchia_likelihood.R:14-27 (without the NA rewrite of removeNAs)."""
    U_obj = createU(vecchia_approx, covparms, nuggets, covmodel, **kw)
    return vecchia_likelihood_U(z, U_obj)


def exact_loglik(z, locs, covparms, nuggets, covmodel="matern"):
    """mvtnorm::dmvnorm(z, sigma = C + diag(nuggets), log=TRUE): the known answer the vignette
    states for m = n-1 (vignettes/GPvecchia_vignette.Rmd:128-139)."""
    locs = np.asarray(locs, dtype=np.float64)
    n = locs.shape[0]
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    Cm = ref_c.MaternFun(D, covparms) if covmodel == "matern" else ref_c.EsqeFun(D, covparms)
    S = Cm + np.diag(np.broadcast_to(np.asarray(nuggets, dtype=np.float64), (n,)))
    L = np.linalg.cholesky(S)
    a = np.linalg.solve(L, np.asarray(z, dtype=np.float64))
    return float(-0.5 * (a @ a) - np.log(np.diag(L)).sum() - 0.5 * n * np.log(2 * np.pi))


# ----------------------------------------------------------------------------------------------
# per-row ("fused") form of the numerator, SURVEY.md 8(a10): what the CUDA kernel accumulates.
# quadform.num = sum_i z_i^2/tau_i + sum_k ( sum_{j: revCond_kj = 0} x_kj * zord[obsrank(id_kj)] )^2
# logdet.num   = -2 sum_k log x_{k,n0} + sum_i log tau_i ; rows < skip_rows (zy dummies) excluded.
# ----------------------------------------------------------------------------------------------
def loglik_numerator_rows(Lentries, revNNarray, revCond, obs, zord, nuggets_ord, row_begin=0,
                          row_end=None, skip_rows=0, include_obs_terms=True):
    revNNarray = np.asarray(revNNarray)
    N, p = revNNarray.shape
    row_end = N if row_end is None else row_end
    obs = np.asarray(obs, dtype=bool)
    obsrank = np.cumsum(obs) - 1
    quad = 0.0
    logd = 0.0
    for k in range(max(row_begin, 0), row_end):
        ids = revNNarray[k]
        keep = ids != NA
        n0 = int(keep.sum())
        if n0 == 0:
            continue
        x = np.asarray(Lentries[k - row_begin])[:n0]
        cond = revCond[k, p - n0:] == 1              # last n0 columns (U_NZentries.cpp:47)
        idk = ids[keep] - 1
        if k < skip_rows:
            continue
        t = 0.0
        for j in range(n0):
            if not cond[j] and obs[idk[j]]:
                t += x[j] * zord[obsrank[idk[j]]]
        quad += t * t
        with np.errstate(divide="ignore"):
            logd += np.log(x[n0 - 1])
    logdet = -2.0 * logd
    if include_obs_terms:
        quad += float(np.sum(np.asarray(zord) ** 2 / np.asarray(nuggets_ord)))
        logdet += float(np.sum(np.log(np.asarray(nuggets_ord))))
    return quad, logdet
