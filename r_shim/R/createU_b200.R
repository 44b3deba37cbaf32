## createU_b200.R -- what changes in the reference's R code (R/createU.R:141-163,
## R/vecchia_likelihood.R:63-76) to use the B200 path.  Signatures and return types are unchanged.
## Not runnable in this repository's image (no R); see INTEGRATION.md.

## device handle cached on the vecchia.approx object (an environment so the cache survives
## pass-by-value); external pointers do not survive saveRDS, so it is re-created lazily
.b200_handle <- function(vecchia.approx) {
  cache <- vecchia.approx$U.prep$b200
  if (is.null(cache)) stop("vecchia_specify() must add U.prep$b200 <- new.env()")
  if (is.null(cache$ptr) || identical(cache$ptr, new("externalptr"))) {
    revNN <- vecchia.approx$U.prep$revNNarray
    cache$ptr <- .Call("_GPvecchia_b200_create", vecchia.approx$locsord, revNN,
                       vecchia.approx$U.prep$revCond, vecchia.approx$obs)
  }
  cache$ptr
}

## replaces R/createU.R:146-162 (the non-MRA, character-covmodel branch)
createU_values_b200 <- function(vecchia.approx, covparms, nuggets.all.ord, nuggets.ord, covmodel,
                                zero.nuggets) {
  h <- .b200_handle(vecchia.approx)
  if (zero.nuggets) {                       # createU.R:83-86 rewrote revCond for this call
    .Call("_GPvecchia_b200_set_revcond", h, vecchia.approx$U.prep$revCond)
    on.exit(.Call("_GPvecchia_b200_set_revcond", h, vecchia.approx$U.prep$revCond.orig))
  }
  ## allLentries = c(c(t(Lentries))[not.na], Zentries), written in that order by the GPU
  allLentries <- .Call("_GPvecchia_b200_U_values", h, covmodel, covparms, nuggets.all.ord, nuggets.ord)
  size <- vecchia.approx$U.prep$size
  Matrix::sparseMatrix(i = vecchia.approx$U.prep$colindices, j = vecchia.approx$U.prep$rowpointers,
                       x = allLentries, dims = c(size, size))
}

## same, without Matrix::sparseMatrix: the library returns the slots of the dgCMatrix (pattern once per
## vecchia.approx, values per call).  Falls back to the triplet route if a conditioning set names the same
## U row twice (the library then reports "unsupported").
createU_csc_b200 <- function(vecchia.approx, covparms, nuggets.all.ord, nuggets.ord, covmodel, zero.nuggets) {
  h <- .b200_handle(vecchia.approx)
  cache <- vecchia.approx$U.prep$b200
  if (is.null(cache$pattern)) cache$pattern <- .Call("_GPvecchia_b200_csc_pattern", h)
  if (zero.nuggets) {
    .Call("_GPvecchia_b200_set_revcond", h, vecchia.approx$U.prep$revCond)
    on.exit(.Call("_GPvecchia_b200_set_revcond", h, vecchia.approx$U.prep$revCond.orig))
  }
  x <- .Call("_GPvecchia_b200_U_values_csc", h, covmodel, covparms, nuggets.all.ord, nuggets.ord)
  size <- as.integer(cache$pattern[[3]])
  methods::new("dgCMatrix", p = cache$pattern[[1]], i = cache$pattern[[2]], x = x, Dim = c(size, size))
}

## optional: numerator of vecchia_likelihood_U (R/vecchia_likelihood.R:74-76) without building U
loglik_numerator_b200 <- function(z, vecchia.approx, covparms, nuggets.all.ord, nuggets.ord, covmodel) {
  h <- .b200_handle(vecchia.approx)
  n <- sum(vecchia.approx$obs)
  skip <- if (vecchia.approx$cond.yz == "zy") n else 0
  r <- .Call("_GPvecchia_b200_loglik_numerator", h, covmodel, covparms, nuggets.all.ord, nuggets.ord,
             z[vecchia.approx$ord.z], skip)
  list(quadform.num = r[1], logdet.num = r[2], nfail = r[3])
}

## the MRA branch (R/createU.R:89-106) needs no R change: the shim registers `_GPvecchia_ic0`,
## `_GPvecchia_createUcppM` and `_GPvecchia_createUcpp` under the reference's own names and arities
## (R/RcppExports.R:53-63), so `createUcpp(ptrs, inds, locsord, covparms)` evaluates the covariances of
## the stored entries on the GPU and returns the incomplete-Cholesky values exactly where
## `Laux = sparseMatrix(j = inds, p = ptrs, x = vals, index1 = FALSE)` (createU.R:109) expects them.
