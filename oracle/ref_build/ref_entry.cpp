// oracle/ref_build/ref_entry.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the UNMODIFIED reference functions compiled from /root/reference/src
// (U_NZentries.cpp, Matern.cpp, Esqe.cpp, dist.cpp, ic0.cpp) against the stand-in headers in
// oracle/ref_build/include.  This file plays the part of the generated glue src/RcppExports.cpp:49-67:
// it turns plain column-major buffers (what R holds) into the argument types the reference functions
// take -- revNNarray by an element-wise converting copy to 64-bit uword, exactly what Rcpp's
// input_parameter<const arma::umat&> does for an R integer matrix -- and copies the returned List out.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
#include <RcppArmadillo.h>
#include <Rcpp.h>
#include <dlfcn.h>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

// the reference's own declarations (src/U_NZentries.cpp:25,126; Matern.h; Esqe.h; ic0.h)
Rcpp::List U_NZentries(const int Ncores, const arma::uword n, const arma::mat& locs, const arma::umat& revNNarray,
                       const arma::mat& revCondOnLatent, const arma::vec& nuggets, const arma::vec& nuggets_obsord,
                       const std::string covType, const arma::vec covparms);
Rcpp::List U_NZentries_mat(int Ncores, const arma::uword n, const arma::mat& locs, const arma::umat& revNNarray,
                           const arma::mat& revCondOnLatent, const arma::vec& nuggets, const arma::vec& nuggets_obsord,
                           arma::mat& covVals, const arma::vec covparms);
arma::mat MaternFun(arma::mat distmat, arma::vec covparms);
arma::mat EsqeFun(arma::mat distmat, arma::vec covparms);
Rcpp::NumericVector ic0(Rcpp::NumericVector ptrs, Rcpp::NumericVector inds, Rcpp::NumericVector vals);
Rcpp::NumericVector createUcppM(Rcpp::NumericVector ptrs, Rcpp::NumericVector inds, Rcpp::NumericVector cov_vals);
Rcpp::NumericVector createUcpp(Rcpp::NumericVector ptrs, Rcpp::NumericVector inds, arma::mat locsord, arma::vec covparams);

namespace {
arma::umat to_umat(const int* cm, long r, long c) {
  arma::umat out((arma::uword)r, (arma::uword)c);
  for (long i = 0; i < r * c; ++i) out.mem[i] = (arma::uword)cm[i];
  return out;
}
void copy_out(const Rcpp::List& res, double* Lentries, double* Zentries) {
  const arma::mat& L = res["Lentries"];
  const arma::mat& Z = res["Zentries"];
  if (L.n_elem) std::memcpy(Lentries, L.memptr(), sizeof(double) * L.n_elem);
  if (Z.n_elem) std::memcpy(Zentries, Z.memptr(), sizeof(double) * Z.n_elem);
}
}  // namespace

extern "C" {

// LAPACK from a shared object (the OpenBLAS inside scipy).  0 = bound.
int gpv_ref_bind_lapack(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  auto p = (arma::lapack_stub::dpotrf_fn)dlsym(h, "scipy_dpotrf_");
  if (!p) p = (arma::lapack_stub::dpotrf_fn)dlsym(h, "dpotrf_");
  auto t = (arma::lapack_stub::dtrtrs_fn)dlsym(h, "scipy_dtrtrs_");
  if (!t) t = (arma::lapack_stub::dtrtrs_fn)dlsym(h, "dtrtrs_");
  if (!p || !t) return 2;
  arma::lapack_stub::g_dpotrf = p;
  arma::lapack_stub::g_dtrtrs = t;
  typedef void (*setnt_fn)(int);
  auto s = (setnt_fn)dlsym(h, "scipy_openblas_set_num_threads");
  if (!s) s = (setnt_fn)dlsym(h, "openblas_set_num_threads");
  if (s) s(1);                                  // 31 x 31 blocks inside an OpenMP loop
  return 0;
}
int gpv_ref_has_lapack() { return arma::lapack_stub::g_dpotrf && arma::lapack_stub::g_dtrtrs; }
// 1: chol/solve by the published unblocked algorithms even when LAPACK is bound (timed beside LAPACK)
void gpv_ref_force_textbook(int on) { arma::lapack_stub::g_force_textbook = on; }
int gpv_ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int gpv_ref_openmp() {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

// .Call('_GPvecchia_U_NZentries', Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord,
//       covType, covparms)   R/RcppExports.R:22-24.   Sizes are explicit here (R reads them off the SEXPs).
// nmsg = number of messages the reference wrote to Rcerr (one per failed Cholesky, :65; one for an unknown
// covType, :28).  Returns 0, or 1 when the reference threw (e.g. unknown covType ends in chol of an empty
// matrix plus an out-of-range span; R would turn that into an error through END_RCPP).
int gpv_ref_U_NZentries(int Ncores, long n, long Nlocs, int d, int p, const double* locs, const int* revNNarray,
                        const double* revCondOnLatent, const double* nuggets, const double* nuggets_obsord,
                        const char* covType, const double* covparms, int ncov, double* Lentries, double* Zentries,
                        long* nmsg) {
  Rcpp::Rcerr.lines = 0;
  int rc = 0;
  try {
    arma::mat locs_m(locs, Nlocs, d);
    arma::umat nn = to_umat(revNNarray, Nlocs, p);
    arma::mat rc_m(revCondOnLatent, Nlocs, p);
    arma::vec nug(nuggets, Nlocs), nug_obs(nuggets_obsord, n), cp(covparms, ncov);
    Rcpp::List res = U_NZentries(Ncores, (arma::uword)n, locs_m, nn, rc_m, nug, nug_obs, std::string(covType), cp);
    copy_out(res, Lentries, Zentries);
  } catch (const std::exception&) {
    rc = 1;
  }
  if (nmsg) *nmsg = Rcpp::Rcerr.lines;
  return rc;
}

// .Call('_GPvecchia_U_NZentries_mat', ...)   R/RcppExports.R:26-28 ; covVals is Nlocs x Nlocs column-major
int gpv_ref_U_NZentries_mat(int Ncores, long n, long Nlocs, int d, int p, const double* locs, const int* revNNarray,
                            const double* revCondOnLatent, const double* nuggets, const double* nuggets_obsord,
                            const double* covVals, const double* covparms, int ncov, double* Lentries,
                            double* Zentries, long* nmsg) {
  Rcpp::Rcerr.lines = 0;
  int rc = 0;
  try {
    arma::mat locs_m(locs, Nlocs, d);
    arma::umat nn = to_umat(revNNarray, Nlocs, p);
    arma::mat rc_m(revCondOnLatent, Nlocs, p);
    arma::vec nug(nuggets, Nlocs), nug_obs(nuggets_obsord, n), cp(covparms, ncov);
    arma::mat cv(covVals, Nlocs, Nlocs);
    Rcpp::List res = U_NZentries_mat(Ncores, (arma::uword)n, locs_m, nn, rc_m, nug, nug_obs, cv, cp);
    copy_out(res, Lentries, Zentries);
  } catch (const std::exception&) {
    rc = 1;
  }
  if (nmsg) *nmsg = Rcpp::Rcerr.lines;
  return rc;
}

// MaternFun(distmat, covparms) / EsqeFun(distmat, covparms): r x c column-major in and out
void gpv_ref_MaternFun(const double* distmat, long r, long c, const double* covparms, double* out) {
  arma::mat res = MaternFun(arma::mat(distmat, r, c), arma::vec(covparms, 3));
  std::memcpy(out, res.memptr(), sizeof(double) * res.n_elem);
}
void gpv_ref_EsqeFun(const double* distmat, long r, long c, const double* covparms, double* out) {
  arma::mat res = EsqeFun(arma::mat(distmat, r, c), arma::vec(covparms, 4));
  std::memcpy(out, res.memptr(), sizeof(double) * res.n_elem);
}

// ic0(ptrs, inds, vals): vals is overwritten in place (the R vector is shared with the C++ handle).
// Returns the number of "ERROR" lines the reference printed (entries right of the diagonal, :57-58).
long gpv_ref_ic0(long N, double* ptrs, double* inds, long nvals, double* vals) {
  Rcpp::Rcout.lines = 0;
  ic0(Rcpp::NumericVector(ptrs, N + 1), Rcpp::NumericVector(inds, nvals), Rcpp::NumericVector(vals, nvals));
  return Rcpp::Rcout.lines;
}
long gpv_ref_createUcppM(long N, double* ptrs, double* inds, long nvals, double* cov_vals) {
  Rcpp::Rcout.lines = 0;
  createUcppM(Rcpp::NumericVector(ptrs, N + 1), Rcpp::NumericVector(inds, nvals), Rcpp::NumericVector(cov_vals, nvals));
  return Rcpp::Rcout.lines;
}
long gpv_ref_createUcpp(long N, int d, double* ptrs, double* inds, long nvals, const double* locsord,
                        const double* covparams, double* out) {
  Rcpp::Rcout.lines = 0;
  Rcpp::NumericVector v = createUcpp(Rcpp::NumericVector(ptrs, N + 1), Rcpp::NumericVector(inds, nvals),
                                     arma::mat(locsord, N, d), arma::vec(covparams, 3));
  for (long i = 0; i < nvals; ++i) out[i] = v[i];
  return Rcpp::Rcout.lines;
}

}  // extern "C"
