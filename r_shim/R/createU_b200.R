## createU_b200.R -- the R side of the drop-in: `createU()` and `vecchia_likelihood()` with the reference's
## signatures and return values (R/createU.R:65, R/vecchia_likelihood.R:14), calling the C-ABI library through the
## `.Call` routines r_shim/src/gpv_shim.c registers.  What changes is createU.R:146-162 (the .Call and the R-level
## assembly of the sparse matrix); the branches this file does not touch are delegated to the reference's own
## function.
##
## Not runnable in this repository's image (no R); the same logic runs in gpvecchia_b200/host.py (`createU`,
## `vecchia_likelihood`) and is what the GPU tests compare with the oracle.  See INTEGRATION.md.
##
## Results: options(GPvecchia.b200.pinned_results = TRUE)   result vectors in page-locked memory (pooled by the
##                                                  shim; 5 ms instead of 13 ms for the 264 MB of n = 1e6, m = 30)
## Devices: options(GPvecchia.b200.device = 0)      one GPU (default 0)
##          options(GPvecchia.b200.devices = 0:7)   one R process, eight GPUs (gpv_multi_*: a worker thread per
##                                                  device inside the library; rows split by contiguous ranges)

## device handle cached on the vecchia.approx object (an environment, so that the cache survives pass-by-value);
## external pointers do not survive saveRDS()/readRDS(): a NULL pointer is re-created lazily
.b200_handle <- function(vecchia.approx) {
  cache <- vecchia.approx$U.prep$b200
  if (is.null(cache)) stop("vecchia_specify() must add U.prep$b200 <- new.env() (INTEGRATION.md)")
  if (is.null(cache$ptr) || identical(cache$ptr, new("externalptr"))) {
    ## revNNarray is passed as it is: the library reads NA (and 0) as "missing", nothing is modified in place
    cache$ptr <- .Call("_GPvecchia_b200_create", vecchia.approx$locsord, vecchia.approx$U.prep$revNNarray,
                       vecchia.approx$U.prep$revCond, vecchia.approx$obs)
    cache$pattern <- NULL
  }
  cache$ptr
}

## dgCMatrix of U from the library's compressed-column output: pattern once per vecchia.approx, values per call.
## Replaces createU.R:146-162 for a character covmodel.  A conditioning set that names the same U row twice has no
## compressed-column form (the library says "unsupported"): then the packed values + sparseMatrix route below.
.b200_U <- function(vecchia.approx, revCond.call, covparms, nuggets.all.ord, nuggets.ord, covmodel) {
  h <- .b200_handle(vecchia.approx)
  cache <- vecchia.approx$U.prep$b200
  size <- vecchia.approx$U.prep$size
  if (!is.null(revCond.call)) {               # createU.R:83-86 changed revCond for this call only
    .Call("_GPvecchia_b200_set_revcond", h, revCond.call)
    on.exit(.Call("_GPvecchia_b200_set_revcond", h, vecchia.approx$U.prep$revCond))
  }
  x <- tryCatch({
    if (is.null(cache$pattern)) cache$pattern <- .Call("_GPvecchia_b200_csc_pattern", h)
    .Call("_GPvecchia_b200_U_values_csc", h, covmodel, covparms, nuggets.all.ord, nuggets.ord)
  }, error = function(e) NULL)
  if (!is.null(x))
    return(methods::new("dgCMatrix", p = cache$pattern[[1]], i = cache$pattern[[2]], x = x, Dim = c(size, size)))
  ## allLentries = c(c(t(Lentries))[not.na], Zentries) (createU.R:158-160), written in that order by the GPU
  allLentries <- .Call("_GPvecchia_b200_U_values", h, covmodel, covparms, nuggets.all.ord, nuggets.ord)
  Matrix::sparseMatrix(i = vecchia.approx$U.prep$colindices, j = vecchia.approx$U.prep$rowpointers,
                       x = allLentries, dims = c(size, size))
}

## the reference's own createU, kept for the cases this file does not change: its calls to U_NZentries(),
## U_NZentries_mat(), createUcpp(), createUcppM() and ic0() resolve to the shim, which registers those routines
## under the reference's own .Call names and arities (src/RcppExports.cpp:155-172)
createU_reference <- get("createU", envir = asNamespace("GPvecchia"))

#' createU with the reference's signature and return value (R/createU.R:65-201).
#' Character covmodel, NN conditioning, no zero nuggets -- the case of vecchia_estimate / vecchia_likelihood /
#' vecchia_prediction at scale -- goes through the device-resident handle: no N x p matrices cross the boundary,
#' U arrives as the slots of a dgCMatrix.  The MRA / ic0 branch (:89-139), a matrix or function covmodel
#' (:149-151) and the zero-nugget bookkeeping (:83-86, :174-193) stay the reference's own R code.
createU <- function(vecchia.approx, covparms, nuggets, covmodel = 'matern') {
  if (vecchia.approx$conditioning == "mra" || !is.character(covmodel) || any(nuggets == 0))
    return(createU_reference(vecchia.approx, covparms, nuggets, covmodel))

  n <- sum(vecchia.approx$obs)
  size <- vecchia.approx$U.prep$size
  latent <- (1:size) %in% vecchia.approx$U.prep$y.ind
  obs <- vecchia.approx$obs
  ord <- vecchia.approx$ord

  if (length(nuggets) == 1 && vecchia.approx$cond.yz != 'zy' && is.null(getOption("GPvecchia.b200.devices"))) {
    ## one scalar crosses the boundary: the device builds nuggets.all.ord / nuggets.ord (the nugget at observed
    ## locations, 0 elsewhere: createU.R:70-78) and the value calls take NULL vectors
    .Call("_GPvecchia_b200_set_scalar_nugget", .b200_handle(vecchia.approx), nuggets)
    U <- .b200_U(vecchia.approx, NULL, covparms, NULL, NULL, covmodel)
  } else {
    ## nuggets per ordered location and per ordered observation (createU.R:74-80)
    nuggets.all <- c(rep_len(nuggets, n), rep(0, sum(latent) - n))
    ord.all <- if (vecchia.approx$cond.yz == 'zy') c(ord[1:n], ord + n) else ord
    U <- .b200_U(vecchia.approx, NULL, covparms, nuggets.all[ord.all], nuggets.all[vecchia.approx$ord.z], covmodel)
  }

  ## response-first ('zy') layouts carry n dummy latent rows/columns (createU.R:166-171)
  if (vecchia.approx$cond.yz == 'zy') {
    dummy <- 2 * seq_len(n) - 1
    U <- U[-dummy, -dummy]
    latent <- latent[-dummy]
    obs <- obs[-(seq_len(n) + n)]
  }
  list(U = U, latent = latent, ord = ord, obs = obs, zero.nugg = list(), ord.pred = vecchia.approx$ord.pred,
       ord.z = vecchia.approx$ord.z, cond.yz = vecchia.approx$cond.yz, ic0 = vecchia.approx$ic0)
}

#' vecchia_likelihood with the reference's signature and return value (R/vecchia_likelihood.R:14-27).
#' For cond.yz = 'z' (every neighbour conditioned on the response) with a character covmodel and no zero
#' nuggets the whole log-likelihood -- numerator AND denominator (:74-91) -- is one fused GPU pass without
#' building U (gpv_loglik_z: U_y U_y^T is diagonal there); every other case builds U with createU() above
#' and evaluates vecchia_likelihood_U() as the reference does.
vecchia_likelihood = function(z, vecchia.approx, covparms, nuggets, covmodel = 'matern') {

  if (vecchia.approx$cond.yz == 'zy')
    warning("cond.yz='zy' will produce a poor likelihood approximation. Use 'SGV' instead.")

  # remove NAs in data and U (vecchia_likelihood.R:20, :45-58): rewrites z and nuggets in this frame
  removeNAs()

  fused = is.character(covmodel) && vecchia.approx$cond.yz == 'z' && vecchia.approx$conditioning != "mra" &&
    all(vecchia.approx$obs) && !any(nuggets == 0)
  if (fused) {
    n = sum(vecchia.approx$obs)
    if (length(nuggets) == 1) nuggets = rep(nuggets, n)
    nuggets.all.ord = nuggets[vecchia.approx$ord]
    nuggets.ord = nuggets[vecchia.approx$ord.z]
    r = tryCatch(.Call("_GPvecchia_b200_loglik_z", .b200_handle(vecchia.approx), covmodel, covparms,
                       nuggets.all.ord, nuggets.ord, z[vecchia.approx$ord.z]), error = function(e) NULL)
    if (!is.null(r)) {
      if (r[6] > 0) warning(sprintf("Cholesky decomposition failed for %d conditioning set(s)", as.integer(r[6])))
      return(r[1])
    }
  }

  # create the U matrix, compute the loglikelihood (vecchia_likelihood.R:23-26)
  U.obj = createU(vecchia.approx, covparms, nuggets, covmodel)
  vecchia_likelihood_U(z, U.obj)
}

## optional: numerator of vecchia_likelihood_U (R/vecchia_likelihood.R:74-76) without building U, any layout
loglik_numerator_b200 <- function(z, vecchia.approx, covparms, nuggets.all.ord, nuggets.ord, covmodel) {
  h <- .b200_handle(vecchia.approx)
  n <- sum(vecchia.approx$obs)
  skip <- if (vecchia.approx$cond.yz == "zy") n else 0
  r <- .Call("_GPvecchia_b200_loglik_numerator", h, covmodel, covparms, nuggets.all.ord, nuggets.ord,
             z[vecchia.approx$ord.z], skip)
  list(quadform.num = r[1], logdet.num = r[2], nfail = r[3])
}

## MaternFun / EsqeFun: the shim registers `_GPvecchia_MaternFun` and `_GPvecchia_EsqeFun` with the reference's
## names and arity (src/RcppExports.cpp:156-157), so the generated stubs R/RcppExports.R:4-10 work unchanged:
##   MaternFun <- function(distmat, covparms) .Call('_GPvecchia_MaternFun', PACKAGE = 'GPvecchia', distmat, covparms)
