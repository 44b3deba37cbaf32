// oracle/u_nzentries_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path
// (GPvecchia 0.1.8).  Nothing under gpvecchia_b200/ (the product) may link,
// import or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker or as
// the reported CPU baseline.
//
// PARITY STATUS: pinned.  The reference's own sources (unmodified) are compiled by oracle/ref_build into
// oracle/_ref/libgpvecchia_ref.so; this restatement reproduces its outputs BIT FOR BIT on the committed
// fixture tests/golden/ref_compiled.npz and on fresh random inputs (tests/test_reference_pin.py).  It is kept
// because it exposes what the compiled reference cannot: a row-range entry point (bench.py samples rows of a
// 1e7-location problem), the textbook and __float128 modes, and no per-row heap traffic in the timed loop.
// Further pins:
//   * the reference's own known-answer test for the three Matern closed forms
//     (tests/testthat/test-MaternFun.r:32-41) -> tests/test_oracle_golden.py
//   * identities the reference states (vignette :128-139, test-createL.r:43-45)
//   * mpmath.besselk golden vectors for the general-nu branch.
//
// Third-party arithmetic the reference reaches through unvendored dependencies:
//   * Armadillo chol(.,"upper") / solve(R, e)  (src/U_NZentries.cpp:61-62)
//       -> LAPACK dpotrf('U') + dtrtrs('U','N','N'); here bound at run time
//       from the OpenBLAS inside scipy (scipy_dpotrf_/scipy_dtrtrs_) via
//       dlopen, with a textbook unblocked fallback (published dpotf2 algorithm).
//   * boost::math::cyl_bessel_k / tgamma (src/Matern.cpp:73,80), BH unpinned
//       -> std::cyl_bessel_k / std::tgamma (the reference itself shipped that
//       substitution in v0.1.5, NEWS.md:11-15).
//
// Layout conventions are R's: all matrices column-major.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <quadmath.h>

extern "C" {

typedef void (*dpotrf_fn)(const char*, const int*, double*, const int*, int*);
typedef void (*dtrtrs_fn)(const char*, const char*, const char*, const int*, const int*,
                          const double*, const int*, double*, const int*, int*);
static dpotrf_fn g_dpotrf = nullptr;
static dtrtrs_fn g_dtrtrs = nullptr;
static void* g_lapack_handle = nullptr;

// Bind LAPACK from a shared object (scipy's bundled OpenBLAS). Returns 0 on success.
int gpv_oracle_bind_lapack(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  const char* potrf_names[] = {"scipy_dpotrf_", "dpotrf_", "scipy_dpotrf_64_", nullptr};
  const char* trtrs_names[] = {"scipy_dtrtrs_", "dtrtrs_", "scipy_dtrtrs_64_", nullptr};
  dpotrf_fn p = nullptr; dtrtrs_fn t = nullptr;
  for (int i = 0; potrf_names[i] && i < 2; ++i) { p = (dpotrf_fn)dlsym(h, potrf_names[i]); if (p) break; }
  for (int i = 0; trtrs_names[i] && i < 2; ++i) { t = (dtrtrs_fn)dlsym(h, trtrs_names[i]); if (t) break; }
  if (!p || !t) { dlclose(h); return 2; }
  g_dpotrf = p; g_dtrtrs = t; g_lapack_handle = h;
  // keep OpenBLAS single-threaded inside the OpenMP loop (31x31 blocks)
  typedef void (*setnt_fn)(int);
  setnt_fn s = (setnt_fn)dlsym(h, "scipy_openblas_set_num_threads");
  if (!s) s = (setnt_fn)dlsym(h, "openblas_set_num_threads");
  if (s) s(1);
  return 0;
}
int gpv_oracle_has_lapack() { return g_dpotrf && g_dtrtrs; }
int gpv_oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

namespace {

// ---- src/dist.cpp:10-16 : Euclidean distance, sequential sum in coordinate order
template <class T>
inline T dist_rows(const T* l1, const T* l2, int d);
template <>
inline double dist_rows<double>(const double* l1, const double* l2, int d) {
  double ssq = 0.0;
  for (int k = 0; k < d; ++k) ssq += (l1[k] - l2[k]) * (l1[k] - l2[k]);
  return std::sqrt(ssq);
}
template <>
inline __float128 dist_rows<__float128>(const __float128* l1, const __float128* l2, int d) {
  __float128 ssq = 0.0;
  for (int k = 0; k < d; ++k) ssq += (l1[k] - l2[k]) * (l1[k] - l2[k]);
  return sqrtq(ssq);
}

// ---- src/dist.cpp:20-30 : full n0 x n0 matrix, no symmetry exploited. x is n0 x d row-major here
// (the gathered rows), out is n0 x n0 column-major.
template <class T>
void calcPWD(const T* x, int n0, int d, T* out) {
  for (int arow = 0; arow < n0; ++arow)
    for (int acol = 0; acol < n0; ++acol)
      out[arow + (size_t)acol * n0] = dist_rows<T>(x + (size_t)arow * d, x + (size_t)acol * d, d);
}

// ---- src/Matern.cpp:24-86
void MaternFun(const double* distmat, int n0, const double* covparms, double* covmat) {
  const int nn = n0 * n0;
  double scaledist;
  if (covparms[2] == 0.5) {                                   // :32-42
    for (int j = 0; j < nn; ++j) {
      if (distmat[j] == 0) covmat[j] = covparms[0];
      else { scaledist = distmat[j] / covparms[1]; covmat[j] = covparms[0] * std::exp(-scaledist); }
    }
  } else if (covparms[2] == 1.5) {                            // :43-57
    for (int j = 0; j < nn; ++j) {
      if (distmat[j] == 0) covmat[j] = covparms[0];
      else {
        scaledist = distmat[j] / covparms[1];
        covmat[j] = covparms[0] * (1 + std::sqrt(3.0) * scaledist) * std::exp(-std::sqrt(3.0) * scaledist);
      }
    }
  } else if (covparms[2] == 2.5) {                            // :58-71
    for (int j = 0; j < nn; ++j) {
      if (distmat[j] == 0) covmat[j] = covparms[0];
      else {
        scaledist = distmat[j] / covparms[1];
        covmat[j] = covparms[0] * std::exp(-scaledist * std::sqrt(5.0)) *
                    (1 + std::sqrt(5.0) * scaledist + 5 * scaledist * scaledist / 3);
      }
    }
  } else {                                                    // :72-83 (no sqrt(2 nu) scaling!)
    double normcon = covparms[0] / (std::pow(2.0, covparms[2] - 1) * std::tgamma(covparms[2]));
    for (int j = 0; j < nn; ++j) {
      if (distmat[j] == 0) covmat[j] = covparms[0];
      else {
        scaledist = distmat[j] / covparms[1];
        covmat[j] = normcon * std::pow(scaledist, covparms[2]) * std::cyl_bessel_k(covparms[2], scaledist);
      }
    }
  }
}

// quad-precision closed forms (arbitration oracle). The general branch has no quad Bessel:
// K_nu is evaluated in double and promoted (documented in the header of the Python wrapper).
void MaternFunQ(const __float128* distmat, int n0, const double* covparms, __float128* covmat) {
  const int nn = n0 * n0;
  const __float128 s2 = covparms[0], range = covparms[1];
  const double nu = covparms[2];
  for (int j = 0; j < nn; ++j) {
    if (distmat[j] == 0) { covmat[j] = s2; continue; }
    __float128 s = distmat[j] / range;
    if (nu == 0.5) covmat[j] = s2 * expq(-s);
    else if (nu == 1.5) covmat[j] = s2 * (1 + sqrtq((__float128)3) * s) * expq(-sqrtq((__float128)3) * s);
    else if (nu == 2.5) covmat[j] = s2 * expq(-s * sqrtq((__float128)5)) * (1 + sqrtq((__float128)5) * s + 5 * s * s / 3);
    else {
      double sd = (double)s;
      double normcon = covparms[0] / (std::pow(2.0, nu - 1) * std::tgamma(nu));
      covmat[j] = (__float128)(normcon * std::pow(sd, nu) * std::cyl_bessel_k(nu, sd));
    }
  }
}

// ---- src/Esqe.cpp:17-39
void EsqeFun(const double* distmat, int n0, const double* covparms, double* covmat) {
  const int nn = n0 * n0;
  for (int j = 0; j < nn; ++j) {
    if (distmat[j] == 0) covmat[j] = covparms[0] + covparms[2];
    else {
      double scaledist = distmat[j] / covparms[1];
      double scaledist2 = std::pow(distmat[j] / covparms[3], 2);
      covmat[j] = covparms[0] * std::exp(-scaledist) + covparms[2] * std::exp(-scaledist2);
    }
  }
}
void EsqeFunQ(const __float128* distmat, int n0, const double* covparms, __float128* covmat) {
  const int nn = n0 * n0;
  for (int j = 0; j < nn; ++j) {
    if (distmat[j] == 0) covmat[j] = (__float128)covparms[0] + covparms[2];
    else {
      __float128 a = distmat[j] / (__float128)covparms[1];
      __float128 b = distmat[j] / (__float128)covparms[3];
      covmat[j] = covparms[0] * expq(-a) + covparms[2] * expq(-b * b);
    }
  }
}

// Textbook unblocked upper Cholesky (LAPACK dpotf2 'U' algorithm as published) + back substitution.
// Used (a) when no LAPACK could be bound, (b) as the quad-precision arbiter.
// Returns 0 on success, j+1 if the leading minor of order j+1 is not positive definite.
template <class T, class SQRT>
int potf2_upper(T* a, int n, SQRT sq) {
  for (int j = 0; j < n; ++j) {
    T ajj = a[j + (size_t)j * n];
    for (int i = 0; i < j; ++i) ajj -= a[i + (size_t)j * n] * a[i + (size_t)j * n];
    if (!(ajj > 0)) return j + 1;  // also catches NaN, like dpotf2's disnan test
    ajj = sq(ajj);
    a[j + (size_t)j * n] = ajj;
    for (int c = j + 1; c < n; ++c) {
      T v = a[j + (size_t)c * n];
      for (int i = 0; i < j; ++i) v -= a[i + (size_t)j * n] * a[i + (size_t)c * n];
      a[j + (size_t)c * n] = v / ajj;
    }
  }
  return 0;
}
template <class T>
void trsv_upper(const T* r, int n, T* b) {  // solves R x = b in place
  for (int i = n - 1; i >= 0; --i) {
    T v = b[i];
    for (int j = i + 1; j < n; ++j) v -= r[i + (size_t)j * n] * b[j];
    b[i] = v / r[i + (size_t)i * n];
  }
}

enum CovKind { COV_MATERN = 0, COV_ESQE = 1, COV_UNKNOWN = 2 };

}  // namespace

extern "C" {

// Exported so tests can pin the covariance functions alone against the reference's
// test-MaternFun.r known answers.  distmat is any array of length len.
void gpv_oracle_MaternFun(const double* distmat, int len_sqrt, const double* covparms, double* out) {
  MaternFun(distmat, len_sqrt, covparms, out);
}
void gpv_oracle_EsqeFun(const double* distmat, int len_sqrt, const double* covparms, double* out) {
  EsqeFun(distmat, len_sqrt, covparms, out);
}
// flat versions (len entries)
void gpv_oracle_MaternFun_flat(const double* distv, long len, const double* covparms, double* out) {
  for (long i = 0; i < len; ++i) MaternFun(distv + i, 1, covparms, out + i);
}
void gpv_oracle_EsqeFun_flat(const double* distv, long len, const double* covparms, double* out) {
  for (long i = 0; i < len; ++i) EsqeFun(distv + i, 1, covparms, out + i);
}

// Restatement of U_NZentries (src/U_NZentries.cpp:25-118), same nine arguments plus explicit sizes
// (R reads the sizes off the SEXPs).
//   Ncores            OpenMP team size (:37)
//   n                 number of observations (length of nuggets_obsord)
//   locs              Nlocs x d, column-major
//   revNNarray        Nlocs x p, column-major, 1-based ids, 0 = missing (createU.R:146-147)
//   revCondOnLatent   Nlocs x p, column-major, double (1 = latent, 0 = response, NaN = missing)
//   nuggets           length Nlocs ; nuggets_obsord length n
//   covType           "matern" | "esqe"
//   Lentries          out, Nlocs x p column-major (zero-initialised here, :33)
//   Zentries          out, length 2n
//   mode              0 = LAPACK (falls back to textbook if unbound), 1 = textbook fp64, 2 = quad arbiter
//   nfail             out, number of rows whose Cholesky failed (left zero, :64-66)
// Returns 0, or 1 for an unknown covType (the reference only prints a message, :27-29, and then
// crashes on an empty covmat; the restatement returns early instead).
// Row-range form used by bench.py's CPU baseline: revNNarray / revCondOnLatent / Lentries hold only
// rows [row_begin, row_begin + nrows) (column-major, leading dimension nrows); locs and nuggets are
// whole.  gpv_oracle_U_NZentries below is the full-range call with the reference's arguments.
int gpv_oracle_U_NZentries_rows(int Ncores, long n, long Nlocs, int d, int p, long row_begin, long nrows,
                                const double* locs, const int* revNNarray, const double* revCondOnLatent,
                                const double* nuggets, const double* nuggets_obsord,
                                const char* covType, const double* covparms,
                                double* Lentries, double* Zentries, int mode, long* nfail);

int gpv_oracle_U_NZentries(int Ncores, long n, long Nlocs, int d, int p,
                           const double* locs, const int* revNNarray, const double* revCondOnLatent,
                           const double* nuggets, const double* nuggets_obsord,
                           const char* covType, const double* covparms,
                           double* Lentries, double* Zentries, int mode, long* nfail) {
  return gpv_oracle_U_NZentries_rows(Ncores, n, Nlocs, d, p, 0, Nlocs, locs, revNNarray, revCondOnLatent,
                                     nuggets, nuggets_obsord, covType, covparms, Lentries, Zentries, mode,
                                     nfail);
}

int gpv_oracle_U_NZentries_rows(int Ncores, long n, long Nlocs, int d, int p, long row_begin, long nrows,
                                const double* locs, const int* revNNarray, const double* revCondOnLatent,
                                const double* nuggets, const double* nuggets_obsord,
                                const char* covType, const double* covparms,
                                double* Lentries, double* Zentries, int mode, long* nfail) {
  (void)row_begin;  // ids inside revNNarray are global; the range only selects which rows are present
  CovKind kind = COV_UNKNOWN;
  if (std::strcmp(covType, "matern") == 0) kind = COV_MATERN;
  else if (std::strcmp(covType, "esqe") == 0) kind = COV_ESQE;
  if (kind == COV_UNKNOWN) {
    std::fprintf(stderr, "Error message: %s covariance is not implemented\n", covType);
    return 1;
  }
  const int m = p - 1;                                       // :31
  std::memset(Lentries, 0, sizeof(double) * (size_t)nrows * p);  // :33
  long fails = 0;
  const bool use_lapack = (mode == 0) && g_dpotrf && g_dtrtrs;
  if (Ncores < 1) Ncores = 1;

#pragma omp parallel for num_threads(Ncores) schedule(static) reduction(+ : fails)
  for (long k = 0; k < nrows; ++k) {                         // :39
    std::vector<long> inds00; inds00.reserve(p);
    for (int j = 0; j < p; ++j) {                            // :41-44  find(inds) - 1
      int id = revNNarray[k + (size_t)j * nrows];
      if (id != 0) inds00.push_back((long)id - 1);
    }
    const int n0 = (int)inds00.size();                       // :45
    if (n0 == 0) continue;
    std::vector<double> nug(n0);                             // :47
    for (int j = 0; j < n0; ++j) {
      double rc = revCondOnLatent[k + (size_t)(m + 1 - n0 + j) * nrows];
      nug[j] = nuggets[inds00[j]] * (1.0 - rc);
    }
    if (mode != 2) {
      std::vector<double> x((size_t)n0 * d), dist((size_t)n0 * n0), covmat((size_t)n0 * n0);
      for (int j = 0; j < n0; ++j)
        for (int c = 0; c < d; ++c) x[(size_t)j * d + c] = locs[inds00[j] + (size_t)c * Nlocs];
      calcPWD<double>(x.data(), n0, d, dist.data());         // :48
      if (kind == COV_MATERN) MaternFun(dist.data(), n0, covparms, covmat.data());  // :51-55
      else EsqeFun(dist.data(), n0, covparms, covmat.data());
      for (int j = 0; j < n0; ++j) covmat[j + (size_t)j * n0] += nug[j];
      std::vector<double> onevec(n0, 0.0);                   // :57-58
      onevec[n0 - 1] = 1;
      int info = 0;
      if (use_lapack) {                                      // :60-62
        g_dpotrf("U", &n0, covmat.data(), &n0, &info);
        if (info == 0) {
          int one = 1, info2 = 0;
          g_dtrtrs("U", "N", "N", &n0, &one, covmat.data(), &n0, onevec.data(), &n0, &info2);
          info = info2;
        }
      } else {
        info = potf2_upper<double>(covmat.data(), n0, [](double v) { return std::sqrt(v); });
        if (info == 0) trsv_upper<double>(covmat.data(), n0, onevec.data());
      }
      if (info == 0) {                                       // :63
        for (int j = 0; j < n0; ++j) Lentries[k + (size_t)j * nrows] = onevec[j];
      } else {
        fails += 1;                                          // :64-66 row stays zero
      }
    } else {
      typedef __float128 Q;
      std::vector<Q> x((size_t)n0 * d), dist((size_t)n0 * n0), covmat((size_t)n0 * n0);
      for (int j = 0; j < n0; ++j)
        for (int c = 0; c < d; ++c) x[(size_t)j * d + c] = locs[inds00[j] + (size_t)c * Nlocs];
      calcPWD<Q>(x.data(), n0, d, dist.data());
      if (kind == COV_MATERN) MaternFunQ(dist.data(), n0, covparms, covmat.data());
      else EsqeFunQ(dist.data(), n0, covparms, covmat.data());
      for (int j = 0; j < n0; ++j) covmat[j + (size_t)j * n0] += nug[j];
      std::vector<Q> onevec(n0, (Q)0);
      onevec[n0 - 1] = 1;
      int info = potf2_upper<Q>(covmat.data(), n0, [](Q v) { return sqrtq(v); });
      if (info == 0) {
        trsv_upper<Q>(covmat.data(), n0, onevec.data());
        for (int j = 0; j < n0; ++j) Lentries[k + (size_t)j * nrows] = (double)onevec[j];
      } else {
        fails += 1;
      }
    }
  }

  for (long i = 0; i < n; ++i) {                             // :110-115
    Zentries[2 * i] = (-1) / std::sqrt(nuggets_obsord[i]);
    Zentries[2 * i + 1] = 1 / std::sqrt(nuggets_obsord[i]);
  }
  if (nfail) *nfail = fails;
  return 0;
}

// Condition number (2-norm estimate via the Cholesky factor is not needed; we report the
// ratio max diag(R)^2 / min diag(R)^2 as a cheap conditioning proxy) for choosing test tolerances.
double gpv_oracle_block_cond_proxy(long k, long Nlocs, int d, int p, const double* locs,
                                   const int* revNNarray, const double* revCondOnLatent,
                                   const double* nuggets, const char* covType, const double* covparms) {
  const int m = p - 1;
  std::vector<long> inds00;
  for (int j = 0; j < p; ++j) {
    int id = revNNarray[k + (size_t)j * Nlocs];
    if (id != 0) inds00.push_back((long)id - 1);
  }
  const int n0 = (int)inds00.size();
  if (n0 == 0) return 0.0;
  std::vector<double> x((size_t)n0 * d), dist((size_t)n0 * n0), covmat((size_t)n0 * n0);
  for (int j = 0; j < n0; ++j)
    for (int c = 0; c < d; ++c) x[(size_t)j * d + c] = locs[inds00[j] + (size_t)c * Nlocs];
  calcPWD<double>(x.data(), n0, d, dist.data());
  if (std::strcmp(covType, "matern") == 0) MaternFun(dist.data(), n0, covparms, covmat.data());
  else EsqeFun(dist.data(), n0, covparms, covmat.data());
  for (int j = 0; j < n0; ++j) {
    double rc = revCondOnLatent[k + (size_t)(m + 1 - n0 + j) * Nlocs];
    covmat[j + (size_t)j * n0] += nuggets[inds00[j]] * (1.0 - rc);
  }
  int info = potf2_upper<double>(covmat.data(), n0, [](double v) { return std::sqrt(v); });
  if (info) return INFINITY;
  double mx = 0, mn = INFINITY;
  for (int j = 0; j < n0; ++j) {
    double r = covmat[j + (size_t)j * n0];
    mx = std::fmax(mx, r); mn = std::fmin(mn, r);
  }
  return (mx / mn) * (mx / mn);
}

// ---- src/ic0.cpp : incomplete Cholesky on a fixed pattern (the MRA branch of createU, createU.R:89-140)
// Restated line by line; indices are doubles because the reference takes Rcpp NumericVectors.
// dot_prod (:15-29): merge of two index ranges [l1, u1] and [l2, u2] (inclusive), both ascending.
static double ic0_dot_prod(long l1, long u1, long l2, long u2, const double* row_inds, const double* cells) {
  double result = 0.0;
  while (l1 <= u1 && l2 <= u2) {
    if (row_inds[l1] == row_inds[l2]) { result += cells[l1] * cells[l2]; l1++; l2++; }
    else if (row_inds[l1] < row_inds[l2]) l1++;
    else l2++;
  }
  return result;
}
// ic0 (:43-63): ptrs has N + 1 entries, vals is overwritten; returns the number of entries right of the
// diagonal (the reference prints "ERROR" for each and leaves the value alone, :57-58).
long gpv_oracle_ic0(long N, const double* ptrs, const double* inds, double* vals) {
  long nerr = 0;
  for (long i = 0; i < N; ++i) {
    for (long j = (long)ptrs[i]; j < (long)ptrs[i + 1]; ++j) {
      const long u1 = (long)ptrs[i];
      const long u2 = (long)ptrs[(long)inds[j]];
      const double dp = ic0_dot_prod(u1, (long)ptrs[i + 1] - 2, u2, (long)ptrs[(long)inds[j] + 1] - 2, inds, vals);
      if (inds[j] < i) vals[j] = (vals[j] - dp) / vals[(long)ptrs[(long)inds[j] + 1] - 1];
      else if (inds[j] == i) vals[j] = std::sqrt(vals[j] - dp);
      else nerr++;
    }
  }
  return nerr;
}
// createUcpp (:77-92): Matern covariance of every (row, stored column) pair, then ic0.  locsord is
// N x d column-major.  fill_only != 0 stops before ic0 (to check the covariance fill by itself).
long gpv_oracle_createUcpp(long N, int d, const double* ptrs, const double* inds, const double* locsord,
                           const double* covparams, int fill_only, double* vals) {
  std::vector<double> a(d), b(d);
  for (long i = 0; i < N; ++i) {
    for (long j = (long)ptrs[i]; j < (long)ptrs[i + 1]; ++j) {
      const long c = (long)inds[j];
      for (int k = 0; k < d; ++k) { a[k] = locsord[i + (size_t)k * N]; b[k] = locsord[c + (size_t)k * N]; }
      const double D = dist_rows<double>(a.data(), b.data(), d);   // :84
      MaternFun(&D, 1, covparams, vals + j);                       // :85
    }
  }
  return fill_only ? 0 : gpv_oracle_ic0(N, ptrs, inds, vals);      // :89
}


}  // extern "C"
