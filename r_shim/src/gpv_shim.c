/* gpv_shim.c -- the `.Call` side of the drop-in boundary, for a GPvecchia maintainer.
 *
 * Registers routines with the SAME argument lists the reference's generated glue has
 * (src/RcppExports.cpp:49-67, R/RcppExports.R:22-24) but without Rcpp/Armadillo: it reads the
 * R vectors in place (INTEGER()/REAL()/LOGICAL()), allocates results with Rf_allocMatrix (owned by
 * R's GC) and calls the C ABI of include/gpvecchia_b200.h.
 *
 * There is no R in this repository's image (no R.h / Rinternals.h): the file is type-checked and RUN against a
 * mock of the R C API (tests/r_api_mock, tests/test_r_shim_mock.py), never inside a real R session.  The same
 * C ABI is exercised by gpvecchia_b200/host.py through ctypes.  Build inside the R package with
 *     PKG_LIBS = -L$(GPV_B200_HOME) -lgpvecchia_b200 -Wl,-rpath,$(GPV_B200_HOME)
 *     PKG_CPPFLAGS = -I$(GPV_B200_HOME)/include
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Rallocators.h>
#include "gpvecchia_b200.h"

static void check(gpv_status st) {
  /* no device work or host allocation is pending here: the C ABI is synchronous and frees its
   * own temporaries before returning, so Rf_error's longjmp is safe. */
  if (st != GPV_OK) Rf_error("gpvecchia_b200: %s", gpv_last_error());
}

static void warn_fail(int64_t nfail, int64_t first_fail) {
  if (nfail > 0)
    Rf_warning("Cholesky decomposition failed for %lld conditioning set(s) (first at row %lld); "
               "those rows of U are zero", (long long)nfail, (long long)first_fail + 1);
}

/* ---- result vectors in page-locked memory (optional: options(GPvecchia.b200.pinned_results = TRUE)) ----------
 * Every ordinary R vector is pageable, and a fresh one faults in page by page: 13 ms for the 264 MB of U at
 * n = 1e6 even with the library's copy workers, 5 ms into page-locked memory.  R lets a package supply the storage
 * of a vector (Rf_allocVector3 + R_allocator_t, R >= 3.1.0): the block comes from gpv_host_alloc and goes back to a
 * small pool when the garbage collector frees the vector, so that the createU() calls of an estimation loop -- same
 * size every time -- reuse it and do not pay for pinning 264 MB again (tens of milliseconds).  The pool holds at
 * most kPinPool blocks; a block that does not fit any more is released. */
#define kPinPool 4
static struct { void* p; size_t cap; int used; } pin_pool[kPinPool];
static void* pin_alloc(R_allocator_t* a, size_t bytes) {
  (void)a;
  int best = -1, victim = -1;
  for (int i = 0; i < kPinPool; ++i) {
    if (pin_pool[i].used) continue;
    if (pin_pool[i].p && pin_pool[i].cap >= bytes && (best < 0 || pin_pool[i].cap < pin_pool[best].cap)) best = i;
    /* where a new block would be remembered: an empty slot, else the smallest idle block */
    if (victim < 0 || (pin_pool[victim].p && (!pin_pool[i].p || pin_pool[i].cap < pin_pool[victim].cap))) victim = i;
  }
  if (best >= 0) { pin_pool[best].used = 1; return pin_pool[best].p; }
  void* p = gpv_host_alloc(bytes);
  if (!p) return NULL;                                       /* R reports the allocation failure */
  if (victim >= 0) {                                         /* remember it for reuse, in place of an idle block */
    if (pin_pool[victim].p) gpv_host_free(pin_pool[victim].p);
    pin_pool[victim].p = p; pin_pool[victim].cap = bytes; pin_pool[victim].used = 1;
  }
  return p;                                                  /* pool full of blocks in use: freed for good on release */
}
static void pin_free(R_allocator_t* a, void* p) {
  (void)a;
  for (int i = 0; i < kPinPool; ++i)
    if (pin_pool[i].p == p) { pin_pool[i].used = 0; return; }
  gpv_host_free(p);
}
static SEXP alloc_result(R_xlen_t n) {
  SEXP opt = Rf_GetOption1(Rf_install("GPvecchia.b200.pinned_results"));
  if (opt != R_NilValue && Rf_asInteger(opt) == 1 && n > 0) {
    R_allocator_t al = {pin_alloc, pin_free, NULL, NULL};    /* R keeps its own copy */
    return Rf_allocVector3(REALSXP, n, &al);
  }
  return Rf_allocVector(REALSXP, n);
}

static int device_from_option(void) {
  SEXP opt = Rf_GetOption1(Rf_install("GPvecchia.b200.device"));
  return (opt == R_NilValue) ? 0 : Rf_asInteger(opt);
}

/* ---- stateless drop-in: same nine arguments, same return value as U_NZentries --------------- */
SEXP gpvb200_U_NZentries(SEXP Ncores, SEXP n, SEXP locs, SEXP revNNarray, SEXP revCondOnLatent,
                         SEXP nuggets, SEXP nuggets_obsord, SEXP covType, SEXP covparms) {
  if (!Rf_isMatrix(locs) || !Rf_isReal(locs)) Rf_error("locs must be a numeric matrix");
  if (!Rf_isMatrix(revNNarray)) Rf_error("revNNarray must be a matrix");
  const int64_t N = Rf_nrows(locs);
  const int d = Rf_ncols(locs), p = Rf_ncols(revNNarray);
  const int64_t nobs = (int64_t)Rf_asReal(n);
  SEXP nn = PROTECT(Rf_coerceVector(revNNarray, INTSXP));    /* createU.R:146-147 already set NA -> 0 */
  const int is_lgl = Rf_isLogical(revCondOnLatent);
  SEXP rc = PROTECT(is_lgl ? revCondOnLatent : Rf_coerceVector(revCondOnLatent, REALSXP));
  SEXP L = PROTECT(Rf_allocMatrix(REALSXP, (int)N, p));
  SEXP Z = PROTECT(Rf_allocMatrix(REALSXP, (int)(2 * nobs), 1));
  int64_t nfail = 0, first = -1;
  gpv_status st = gpv_U_NZentries(Rf_asInteger(Ncores), nobs, N, p, d, REAL(locs), INTEGER(nn),
                                  is_lgl ? (const void*)LOGICAL(rc) : (const void*)REAL(rc),
                                  is_lgl ? GPV_COND_RLOGICAL_I32 : GPV_COND_F64, REAL(nuggets),
                                  REAL(nuggets_obsord), CHAR(STRING_ELT(covType, 0)), REAL(covparms),
                                  LENGTH(covparms), REAL(L), REAL(Z), &nfail, &first,
                                  device_from_option());
  check(st);
  warn_fail(nfail, first);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, L);
  SET_VECTOR_ELT(out, 1, Z);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("Lentries"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("Zentries"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(6);
  return out;
}

/* ---- stateless drop-in for the matrix branch: same nine arguments as U_NZentries_mat
 * (src/RcppExports.cpp: _GPvecchia_U_NZentries_mat; covVals takes the place of covType) -------------- */
SEXP gpvb200_U_NZentries_mat(SEXP Ncores, SEXP n, SEXP locs, SEXP revNNarray, SEXP revCondOnLatent,
                             SEXP nuggets, SEXP nuggets_obsord, SEXP covVals, SEXP covparms) {
  (void)Ncores; (void)nuggets; (void)covparms;               /* unused by the reference too (:126-197) */
  const int64_t N = Rf_nrows(locs);
  const int d = Rf_ncols(locs), p = Rf_ncols(revNNarray);
  const int64_t nobs = (int64_t)Rf_asReal(n);
  if (!Rf_isReal(covVals) || Rf_nrows(covVals) != N || Rf_ncols(covVals) != N) Rf_error("covmodel must be an N x N numeric matrix");
  SEXP nn = PROTECT(Rf_coerceVector(revNNarray, INTSXP));
  const int is_lgl = Rf_isLogical(revCondOnLatent);
  SEXP rc = PROTECT(is_lgl ? revCondOnLatent : Rf_coerceVector(revCondOnLatent, REALSXP));
  SEXP L = PROTECT(Rf_allocMatrix(REALSXP, (int)N, p));
  SEXP Z = PROTECT(Rf_allocMatrix(REALSXP, (int)(2 * nobs), 1));
  gpv_handle* h = NULL;
  check(gpv_create(&h, N, p, d, REAL(locs), INTEGER(nn), is_lgl ? (const void*)LOGICAL(rc) : (const void*)REAL(rc),
                   is_lgl ? GPV_COND_RLOGICAL_I32 : GPV_COND_F64, NULL, 0, N, device_from_option()));
  int64_t nfail = 0, first = -1;
  gpv_status st = gpv_u_nzentries_mat(h, REAL(covVals), REAL(nuggets_obsord), nobs, REAL(L), REAL(Z), &nfail, &first);
  gpv_destroy(h);
  check(st);
  warn_fail(nfail, first);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, L);
  SET_VECTOR_ELT(out, 1, Z);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(nm, 0, Rf_mkChar("Lentries"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("Zentries"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(6);
  return out;
}

/* ---- device-resident handle: one per vecchia.approx ------------------------------------------ */
static void handle_finalizer(SEXP ptr) {
  gpv_handle* h = (gpv_handle*)R_ExternalPtrAddr(ptr);
  if (h) { gpv_destroy(h); R_ClearExternalPtr(ptr); }
}
static void multi_finalizer(SEXP ptr) {
  gpv_multi* m = (gpv_multi*)R_ExternalPtrAddr(ptr);
  if (m) { gpv_multi_destroy(m); R_ClearExternalPtr(ptr); }
}
static int is_multi(SEXP ptr) { return R_ExternalPtrTag(ptr) == Rf_install("gpv_multi"); }
static gpv_multi* get_multi(SEXP ptr) {
  gpv_multi* m = (gpv_multi*)R_ExternalPtrAddr(ptr);
  if (!m) Rf_error("gpvecchia_b200: stale device handle (vecchia.approx was deserialised?); re-create it");
  return m;
}

SEXP gpvb200_create(SEXP locsord, SEXP revNNarray, SEXP revCond, SEXP obs) {
  const int64_t N = Rf_nrows(locsord);
  const int d = Rf_ncols(locsord), p = Rf_ncols(revNNarray);
  /* Rf_coerceVector returns its ARGUMENT when the type already matches (the usual case: GpGp / FNN neighbour
   * arrays are integer), so nothing may be written through `nn`: is.na(vecchia.approx$U.prep$revNNarray) is
   * still needed by the reference's R code after this call (createU.R:17-18, 90-91, 158).  The library reads
   * NA_integer_ (INT_MIN), like 0, as "missing" (prep_nn_kernel), so createU.R:146-147 needs no copy here. */
  SEXP nn = PROTECT(Rf_coerceVector(revNNarray, INTSXP));
  const int* ip = INTEGER(nn);
  SEXP ob = PROTECT(Rf_coerceVector(obs, LGLSXP));
  SEXP devs = Rf_GetOption1(Rf_install("GPvecchia.b200.devices"));   /* e.g. options(GPvecchia.b200.devices = 0:7) */
  SEXP ptr;
  if (devs != R_NilValue && XLENGTH(devs) > 1) {
    /* one R process, several GPUs: rows split into contiguous ranges, one worker thread per device inside the
     * library (gpv_multi_*); the handle is tagged so the calls below dispatch on it */
    SEXP di = PROTECT(Rf_coerceVector(devs, INTSXP));
    gpv_multi* m = NULL;
    gpv_status st = gpv_multi_create(&m, N, p, d, REAL(locsord), ip, LOGICAL(revCond), GPV_COND_RLOGICAL_I32,
                                     LOGICAL(ob), INTEGER(di), (int)XLENGTH(di));
    UNPROTECT(1);
    check(st);
    ptr = PROTECT(R_MakeExternalPtr(m, Rf_install("gpv_multi"), R_NilValue));
    R_RegisterCFinalizerEx(ptr, multi_finalizer, TRUE);
  } else {
    gpv_handle* h = NULL;
    check(gpv_create(&h, N, p, d, REAL(locsord), ip, LOGICAL(revCond), GPV_COND_RLOGICAL_I32,
                     LOGICAL(ob), 0, N, devs != R_NilValue ? Rf_asInteger(devs) : device_from_option()));
    ptr = PROTECT(R_MakeExternalPtr(h, Rf_install("gpv_handle"), R_NilValue));
    R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  }
  double nobs = 0;                                            /* for likelihood calls that pass zord = NULL */
  for (R_xlen_t i = 0; i < XLENGTH(ob); ++i) nobs += (LOGICAL(ob)[i] == TRUE);
  Rf_setAttrib(ptr, Rf_install("n_obs"), Rf_ScalarReal(nobs));
  UNPROTECT(3);
  return ptr;
}

static gpv_handle* get_handle(SEXP ptr) {
  gpv_handle* h = (gpv_handle*)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("gpvecchia_b200: stale device handle (vecchia.approx was deserialised?); re-create it");
  return h;
}

SEXP gpvb200_set_revcond(SEXP ptr, SEXP revCond) {
  if (is_multi(ptr)) check(gpv_multi_set_revcond(get_multi(ptr), LOGICAL(revCond), GPV_COND_RLOGICAL_I32));
  else check(gpv_set_revcond(get_handle(ptr), LOGICAL(revCond), GPV_COND_RLOGICAL_I32));
  return R_NilValue;
}

/* allLentries of createU.R:158-160 (packed U values followed by Zentries), straight from the GPU */
SEXP gpvb200_U_values(SEXP ptr, SEXP covType, SEXP covparms, SEXP nuggets_all_ord, SEXP nuggets_ord) {
  const int multi = is_multi(ptr);
  /* NULL nuggets (single-device handle): those of _GPvecchia_b200_set_scalar_nugget, already on the device */
  const int resident = Rf_isNull(nuggets_all_ord) && Rf_isNull(nuggets_ord);
  if (resident && multi) Rf_error("resident scalar nugget: single-device handle only");
  const int64_t n = resident ? (int64_t)Rf_asReal(Rf_getAttrib(ptr, Rf_install("n_obs"))) : XLENGTH(nuggets_ord);
  const int64_t len = multi ? gpv_multi_packed_len(get_multi(ptr)) : gpv_packed_len(get_handle(ptr));
  SEXP out = PROTECT(alloc_result((R_xlen_t)(len + 2 * n)));
  int64_t nfail = 0, first = -1;
  gpv_status st = multi
      ? gpv_multi_u_values_packed(get_multi(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                                  REAL(nuggets_all_ord), REAL(nuggets_ord), n, 1, REAL(out), &nfail, &first)
      : gpv_u_values_packed(get_handle(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                            resident ? NULL : REAL(nuggets_all_ord), resident ? NULL : REAL(nuggets_ord), n, 1,
                            REAL(out), &nfail, &first);
  if (st != GPV_OK) { UNPROTECT(1); check(st); }
  warn_fail(nfail, first);
  UNPROTECT(1);
  return out;
}

/* dgCMatrix slots of U: list(p, i, size) once per vecchia.approx ... */
SEXP gpvb200_csc_pattern(SEXP ptr) {
  const int multi = is_multi(ptr);
  int64_t ncols = 0, nnz = 0, size = 0;
  check(multi ? gpv_multi_csc_dims(get_multi(ptr), &ncols, &nnz, &size) : gpv_csc_dims(get_handle(ptr), &ncols, &nnz, &size));
  SEXP p = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)(ncols + 1)));
  SEXP i = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
  gpv_status st = multi ? gpv_multi_u_csc_pattern(get_multi(ptr), INTEGER(p), INTEGER(i))
                        : gpv_u_csc_pattern(get_handle(ptr), INTEGER(p), INTEGER(i));
  if (st != GPV_OK) { UNPROTECT(2); check(st); }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, p);
  SET_VECTOR_ELT(out, 1, i);
  SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)size));
  UNPROTECT(3);
  return out;
}
/* ... and @x per createU call: kernel + createU.R:158-160 + the triplet sort of sparseMatrix (:161) */
SEXP gpvb200_U_values_csc(SEXP ptr, SEXP covType, SEXP covparms, SEXP nuggets_all_ord, SEXP nuggets_ord) {
  const int multi = is_multi(ptr);
  const int resident = Rf_isNull(nuggets_all_ord) && Rf_isNull(nuggets_ord);   /* see gpvb200_U_values */
  if (resident && multi) Rf_error("resident scalar nugget: single-device handle only");
  const int64_t n = resident ? (int64_t)Rf_asReal(Rf_getAttrib(ptr, Rf_install("n_obs"))) : XLENGTH(nuggets_ord);
  int64_t nnz = 0, nfail = 0, first = -1;
  check(multi ? gpv_multi_csc_dims(get_multi(ptr), NULL, &nnz, NULL) : gpv_csc_dims(get_handle(ptr), NULL, &nnz, NULL));
  SEXP out = PROTECT(alloc_result((R_xlen_t)nnz));
  gpv_status st = multi
      ? gpv_multi_u_values_csc(get_multi(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                               REAL(nuggets_all_ord), REAL(nuggets_ord), n, REAL(out), &nfail, &first)
      : gpv_u_values_csc(get_handle(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                         resident ? NULL : REAL(nuggets_all_ord), resident ? NULL : REAL(nuggets_ord), n, REAL(out),
                         &nfail, &first);
  if (st != GPV_OK) { UNPROTECT(1); check(st); }
  warn_fail(nfail, first);
  UNPROTECT(1);
  return out;
}
/* list(colindices, rowpointers) of R/U_sparsity.R:36-73 */
SEXP gpvb200_U_sparsity(SEXP ptr) {
  if (is_multi(ptr)) Rf_error("U_sparsity arrays: use a single-device handle (or the compressed-column pattern)");
  gpv_handle* h = get_handle(ptr);
  int64_t nnz = 0;
  check(gpv_csc_dims(h, NULL, &nnz, NULL));
  SEXP ci = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
  SEXP rp = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
  check(gpv_u_sparsity(h, INTEGER(ci), INTEGER(rp)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, ci);
  SET_VECTOR_ELT(out, 1, rp);
  UNPROTECT(3);
  return out;
}

/* scalar nugget built on the device; afterwards the likelihood and U-values calls may pass NULL (R: NULL) vectors */
SEXP gpvb200_set_scalar_nugget(SEXP ptr, SEXP nugget) {
  if (is_multi(ptr)) Rf_error("resident scalar nugget: single-device handle only");
  check(gpv_set_scalar_nugget(get_handle(ptr), Rf_asReal(nugget)));
  return R_NilValue;
}

/* c(quadform.num, logdet.num, nfail) of vecchia_likelihood.R:74-76, no U materialisation */
SEXP gpvb200_loglik_numerator(SEXP ptr, SEXP covType, SEXP covparms, SEXP nuggets_all_ord,
                              SEXP nuggets_ord, SEXP zord, SEXP skip_rows) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 3));
  if (is_multi(ptr)) {
    if (Rf_isNull(nuggets_all_ord) || Rf_isNull(nuggets_ord) || Rf_isNull(zord)) { UNPROTECT(1); Rf_error("the multi-device handle takes all three vectors"); }
    gpv_status st = gpv_multi_loglik_numerator(get_multi(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                                               REAL(nuggets_all_ord), REAL(nuggets_ord), REAL(zord), XLENGTH(zord),
                                               (int64_t)Rf_asReal(skip_rows), REAL(out));
    UNPROTECT(1);
    check(st);
    return out;
  }
  gpv_handle* h = get_handle(ptr);
  check(gpv_loglik_numerator(h, CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                             Rf_isNull(nuggets_all_ord) ? NULL : REAL(nuggets_all_ord),
                             Rf_isNull(nuggets_ord) ? NULL : REAL(nuggets_ord), Rf_isNull(zord) ? NULL : REAL(zord),
                             Rf_isNull(zord) ? (int64_t)Rf_asReal(Rf_getAttrib(ptr, Rf_install("n_obs"))) : XLENGTH(zord),
                             (int64_t)Rf_asReal(skip_rows), -1, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* c(loglik, quadform.num, logdet.num, quadform.denom, logdet.denom, nfail): the whole vecchia_likelihood for
 * cond.yz = 'z' (every neighbour conditioned on the response: U_y U_y^T is diagonal, so the denominator of
 * vecchia_likelihood.R:85-91 is a per-row closed form); other layouts return an error and R falls back to U2V */
SEXP gpvb200_loglik_z(SEXP ptr, SEXP covType, SEXP covparms, SEXP nuggets_all_ord, SEXP nuggets_ord, SEXP zord) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 6));
  gpv_status st;
  if (is_multi(ptr))
    st = gpv_multi_loglik_z(get_multi(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                            REAL(nuggets_all_ord), REAL(nuggets_ord), REAL(zord), XLENGTH(zord), REAL(out));
  else
    st = gpv_loglik_z(get_handle(ptr), CHAR(STRING_ELT(covType, 0)), REAL(covparms), LENGTH(covparms),
                      Rf_isNull(nuggets_all_ord) ? NULL : REAL(nuggets_all_ord),
                      Rf_isNull(nuggets_ord) ? NULL : REAL(nuggets_ord), Rf_isNull(zord) ? NULL : REAL(zord),
                      Rf_isNull(zord) ? (int64_t)Rf_asReal(Rf_getAttrib(ptr, Rf_install("n_obs"))) : XLENGTH(zord), -1, REAL(out));
  UNPROTECT(1);
  check(st);
  return out;
}

/* MaternFun(distmat, covparms) / EsqeFun(distmat, covparms): the reference exports MaternFun to users
 * (NAMESPACE:3) and has a .Call wrapper for EsqeFun (R/RcppExports.R:4-10); same names, same shapes */
SEXP gpvb200_MaternFun(SEXP distmat, SEXP covparms) {
  if (!Rf_isReal(distmat) || !Rf_isReal(covparms) || XLENGTH(covparms) < 3) Rf_error("MaternFun(distmat, c(sig2, range, smooth))");
  SEXP out = PROTECT(Rf_duplicate(distmat));                 /* keeps dim */
  gpv_status st = gpv_MaternFun(REAL(distmat), XLENGTH(distmat), REAL(covparms), REAL(out), device_from_option());
  UNPROTECT(1);
  check(st);
  return out;
}
SEXP gpvb200_EsqeFun(SEXP distmat, SEXP covparms) {
  if (!Rf_isReal(distmat) || !Rf_isReal(covparms) || XLENGTH(covparms) < 4) Rf_error("EsqeFun(distmat, c(sig2_1, r1, sig2_2, r2))");
  SEXP out = PROTECT(Rf_duplicate(distmat));
  gpv_status st = gpv_EsqeFun(REAL(distmat), XLENGTH(distmat), REAL(covparms), REAL(out), device_from_option());
  UNPROTECT(1);
  check(st);
  return out;
}

/* ---- ic0 / createUcppM / createUcpp: same names, arguments and aliasing as src/ic0.cpp:43-92 -------
 * The reference's NumericVector arguments wrap the R vectors without a copy, so ic0 overwrites `vals` in
 * place and returns it (createUcppM depends on that, :69-70); the shim keeps that behaviour for a
 * double vector and works on a coerced copy otherwise, as Rcpp would. */
static SEXP as_real(SEXP x, int* nprot) {
  if (TYPEOF(x) == REALSXP) return x;
  SEXP y = PROTECT(Rf_coerceVector(x, REALSXP));
  (*nprot)++;
  return y;
}
SEXP gpvb200_ic0(SEXP ptrs, SEXP inds, SEXP vals) {
  int np = 0;
  SEXP p = as_real(ptrs, &np), i = as_real(inds, &np), v = as_real(vals, &np);
  if (XLENGTH(v) != XLENGTH(i)) { UNPROTECT(np); Rf_error("ic0: vals and inds differ in length"); }
  gpv_status st = gpv_ic0((int64_t)XLENGTH(p) - 1, REAL(p), REAL(i), (int64_t)XLENGTH(i), REAL(v));
  UNPROTECT(np);
  check(st);
  return v;
}
SEXP gpvb200_createUcppM(SEXP ptrs, SEXP inds, SEXP cov_vals) { return gpvb200_ic0(ptrs, inds, cov_vals); }
SEXP gpvb200_createUcpp(SEXP ptrs, SEXP inds, SEXP locsord, SEXP covparams) {
  int np = 0;
  SEXP p = as_real(ptrs, &np), i = as_real(inds, &np), l = as_real(locsord, &np), c = as_real(covparams, &np);
  const int64_t N = (int64_t)XLENGTH(p) - 1;
  if (!Rf_isMatrix(l) || Rf_nrows(l) != N || XLENGTH(c) < 3) { UNPROTECT(np); Rf_error("createUcpp: locsord must be an N x d matrix and covparams = c(sig2, range, smooth)"); }
  SEXP vals = PROTECT(Rf_allocVector(REALSXP, XLENGTH(i)));
  np++;
  gpv_status st = gpv_createUcpp(N, Rf_ncols(l), REAL(p), REAL(i), (int64_t)XLENGTH(i), REAL(l), REAL(c), REAL(vals),
                                 device_from_option());
  UNPROTECT(np);
  check(st);
  return vals;
}

static const R_CallMethodDef CallEntries[] = {
    {"_GPvecchia_U_NZentries", (DL_FUNC)&gpvb200_U_NZentries, 9},   /* same name and arity as
                                                                       src/RcppExports.cpp:159 */
    {"_GPvecchia_U_NZentries_mat", (DL_FUNC)&gpvb200_U_NZentries_mat, 9},
    {"_GPvecchia_b200_create", (DL_FUNC)&gpvb200_create, 4},
    {"_GPvecchia_b200_set_revcond", (DL_FUNC)&gpvb200_set_revcond, 2},
    {"_GPvecchia_b200_U_values", (DL_FUNC)&gpvb200_U_values, 5},
    {"_GPvecchia_b200_csc_pattern", (DL_FUNC)&gpvb200_csc_pattern, 1},
    {"_GPvecchia_b200_U_values_csc", (DL_FUNC)&gpvb200_U_values_csc, 5},
    {"_GPvecchia_b200_U_sparsity", (DL_FUNC)&gpvb200_U_sparsity, 1},
    {"_GPvecchia_b200_set_scalar_nugget", (DL_FUNC)&gpvb200_set_scalar_nugget, 2},
    {"_GPvecchia_b200_loglik_numerator", (DL_FUNC)&gpvb200_loglik_numerator, 7},
    {"_GPvecchia_b200_loglik_z", (DL_FUNC)&gpvb200_loglik_z, 6},
    {"_GPvecchia_MaternFun", (DL_FUNC)&gpvb200_MaternFun, 2},      /* src/RcppExports.cpp:156-157: same names, arity 2 */
    {"_GPvecchia_EsqeFun", (DL_FUNC)&gpvb200_EsqeFun, 2},
    {"_GPvecchia_ic0", (DL_FUNC)&gpvb200_ic0, 3},                   /* src/RcppExports.cpp: same names and arities */
    {"_GPvecchia_createUcppM", (DL_FUNC)&gpvb200_createUcppM, 3},
    {"_GPvecchia_createUcpp", (DL_FUNC)&gpvb200_createUcpp, 4},
    {NULL, NULL, 0}};

void R_unload_GPvecchiaB200(DllInfo* dll) {
  (void)dll;
  gpv_release_cached();                       /* the handle the stateless U_NZentries route keeps between calls */
  for (int i = 0; i < kPinPool; ++i)
    if (pin_pool[i].p && !pin_pool[i].used) { gpv_host_free(pin_pool[i].p); pin_pool[i].p = NULL; pin_pool[i].cap = 0; }
}

void R_init_GPvecchiaB200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
