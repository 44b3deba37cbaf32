"""CPU oracle for the createU / U_NZentries / likelihood-numerator path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``gpvecchia_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, as the checker or as the reported CPU baseline.

Parity status: the reference (GPvecchia 0.1.8, R + Rcpp + RcppArmadillo + BH) cannot be
built in this image, so this is a *restatement*.  It is pinned by the reference's only
known-answer test on this path (Matern closed forms, tests/testthat/test-MaternFun.r) and
by the identities the reference states (exact log-density for m = n-1).  Cholesky/solve and
NN-path U values are **parity unpinned** by the reference's own tests.  The ic0
restatement (src/ic0.cpp) is pinned by tests/testthat/test-createL.r:43-45 (full pattern:
L L^T = Sigma to 1e-10 on the 20 x 20 grid).
"""
from .ref_c import (MaternFun, EsqeFun, U_NZentries, block_cond_proxy, lib, max_threads,
                    has_lapack, RowsProblem, ic0, createUcpp, createUcppM)
from .vecchia_np import (U_sparsity, createU, vecchia_likelihood, vecchia_likelihood_U, U2V,
                         vecchia_specify, whichCondOnLatent, find_ordered_nn_brute,
                         loglik_numerator_from_U, exact_loglik, loglik_numerator_rows, U_NZentries_mat)

__all__ = [
    "MaternFun", "EsqeFun", "U_NZentries", "block_cond_proxy", "lib", "max_threads", "has_lapack", "RowsProblem", "ic0", "createUcpp", "createUcppM",
    "U_sparsity", "createU", "vecchia_likelihood", "vecchia_likelihood_U", "U2V",
    "vecchia_specify", "whichCondOnLatent", "find_ordered_nn_brute",
    "loglik_numerator_from_U", "exact_loglik", "loglik_numerator_rows", "U_NZentries_mat",
]
