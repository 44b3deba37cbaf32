"""gpvecchia_b200 -- B200-native (sm_100a) implementation of GPvecchia's createU / U_NZentries
hot path behind the reference's own interface.  The CUDA C-ABI library
(gpvecchia_b200/libgpvecchia_b200.so, include/gpvecchia_b200.h) is the product; this package is
the host-side mirror of the reference's R interface for that path.  Importing fails loudly when
the library has not been built: there is no CPU fallback."""
from ._lib import lib, GpvError, LIB_PATH
from .host import (UHandle, MultiHandle, U_NZentries, MaternFun, EsqeFun, U_sparsity, createU,
                   vecchia_likelihood, vecchia_likelihood_U, vecchia_loglik_numerator,
                   ic0, createUcpp, createUcppM)

__all__ = ["lib", "GpvError", "LIB_PATH", "UHandle", "MultiHandle", "U_NZentries", "MaternFun", "EsqeFun",
           "U_sparsity", "createU", "vecchia_likelihood", "vecchia_likelihood_U",
           "vecchia_loglik_numerator", "ic0", "createUcpp", "createUcppM"]
