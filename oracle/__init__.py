"""CPU oracle for the createU / U_NZentries / likelihood-numerator path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``gpvecchia_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, as the checker or as the reported CPU baseline.

Parity status: PINNED by reference-run outputs.  The reference's own hot-path sources
(/root/reference/src/{U_NZentries,Matern,Esqe,dist,ic0}.cpp, unmodified) are compiled by
oracle/ref_build (stand-in headers for the absent Armadillo / Rcpp / Boost, LAPACK from scipy's
OpenBLAS) into oracle/_ref/libgpvecchia_ref.so (``oracle.ref_native``); its outputs on the inputs of
tests/ref_cases.py are committed as tests/golden/ref_compiled.npz, and this restatement reproduces
them BIT FOR BIT (tests/test_reference_pin.py), as well as the compiled reference run live on fresh
random inputs.  Also pinned by the reference's known answers (Matern closed forms,
tests/testthat/test-MaternFun.r; L L^T = Sigma, test-createL.r:43-45), mpmath golden vectors and the
exact-density identities.  What stays outside: Armadillo's / Boost's own rounding (unpinned versions).
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
