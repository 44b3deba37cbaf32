// bessel_table.cuh -- general-nu Matern branch (src/Matern.cpp:72-83):
//     cov(d) = sig2 / (2^(nu-1) Gamma(nu)) * s^nu * K_nu(s),  s = d / range   (no sqrt(2 nu)!)
//
// The reference evaluates boost::math::cyl_bessel_k per matrix entry (Temme series for s <= 2,
// Steed's CF2 for s > 2, forward recurrence in the order).  nu is constant within one call, so the
// B200 path does that work ONCE per call for a table, not once per pair:
//   * `bessel_k_general` : the published Temme / CF2 algorithm (Temme 1975; Numerical Recipes
//     6.7 `bessik`; also what libstdc++'s std::cyl_bessel_k implements), host+device, used (a) by
//     the table builder and (b) as the in-kernel fallback for arguments outside the table;
//   * `build_cov_table_kernel` : per call, fits cov as a function of w = d^2 by degree-kTabDeg
//     Chebyshev interpolation on intervals that are geometric in w (octave x 2^kTabSubBits), so
//     the branch point at w = 0 is always a fixed relative distance away (uniform accuracy for any
//     nu); for s >= s_split the table stores exp(+s) cov (smooth), and the kernel multiplies exp(-s)
//     back;
//   * `cov_general` : interval index straight from the exponent/mantissa bits of w (no sqrt, no
//     exp, no pow for s < s_split), Horner over coefficient-major rows.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "u_kernels.cuh"

namespace gpv {

#ifndef GPV_TAB_DEG
#define GPV_TAB_DEG 10
#endif
#ifndef GPV_TAB_SUBBITS
#define GPV_TAB_SUBBITS 2
#endif
#ifndef GPV_TAB_WIN_OCT
#define GPV_TAB_WIN_OCT 24      // octaves of w in the shared-memory window of the band kernels
#endif
#ifndef GPV_TAB_WIN_TOP
#define GPV_TAB_WIN_TOP 8       // the window ends this many octaves above the median squared neighbour distance
#endif
constexpr int kTabDeg = GPV_TAB_DEG;
constexpr int kTabSubBits = GPV_TAB_SUBBITS;   // 2^bits intervals per octave of w
constexpr int kTabSub = 1 << kTabSubBits;
constexpr int kTabOctaves = 64;             // covered range of w below its maximum
constexpr double kTabSSplit = 16.0;         // s >= split: table holds exp(s) * cov
constexpr int kTabStride = kTabOctaves * kTabSub;   // intervals per coefficient row
// Coefficient-major storage: coefficient k of interval i sits at coef[k * kTabStride + i] (every
// offset an immediate).  The pair stage of the general branch is bound by these gathers (L1 / shared data pipe
// 85-87 % busy, 2.2-2.4 wavefronts per 8-byte gather, profiles/r01_*, r02_*), so what counts is the NUMBER of
// coefficients.  Round 1 used (sub_bits, degree) = (1, 19); measured against mpmath, the error that matters for
// a covariance MATRIX -- absolute, relative to sigma^2 -- is already at the rounding floor of the Horner
// evaluation (2e-15) with (2, 10), (3, 8) or (4, 7), for any nu in [0.05, 10] and s up to the exp split
// (tools/tab_degree_scan.py); relative accuracy holds to 3e-15 for s <= 3 and degrades as exp(s) beyond, where
// the covariance itself is below 0.05 sigma^2.  (2, 10): 11 gathers instead of 20, and 24 octaves of w fit the
// 8.4 KB shared-memory window of the band kernels.
__host__ __device__ constexpr int tab_coef_index(int k, int i) { return k * kTabStride + i; }

__host__ __device__ inline int hi32_of(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
__host__ __device__ inline int lo32_of(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}
__host__ __device__ inline double from_hilo(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x;
#endif
}

// K_nu(x) (scaled = false) or exp(x) K_nu(x) (scaled = true), nu = t.nl + t.xmu, |xmu| <= 1/2.
__host__ __device__ inline double bessel_k_general(double x, const CovTable& t, bool scaled) {
  const double kEps = 1.0e-16;
  const double kPi = 3.141592653589793238462643383279502884;
  const double xmu = t.xmu, xmu2 = xmu * xmu;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  double rkmu, rk1;
  if (x < 2.0) {
    // Temme's series for K_mu, K_{mu+1}
    const double b = 0.5 * x;
    double d = -log(b);
    double e = xmu * d;
    const double fact2 = (fabs(e) < kEps) ? 1.0 : sinh(e) / e;
    const double pimu = kPi * xmu;
    const double fact = (fabs(pimu) < kEps) ? 1.0 : pimu / sin(pimu);
    double ff = fact * (t.gam1 * cosh(e) + t.gam2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / t.gampl;
    double q = 0.5 / (e * t.gammi);
    double c = 1.0;
    d = b * b;
    double sum1 = p;
    for (int i = 1; i <= 500; ++i) {
      ff = (i * ff + p + q) / (i * (double)i - xmu2);
      c *= (d / i);
      p /= (i - xmu);
      q /= (i + xmu);
      const double del = c * ff;
      sum += del;
      const double del1 = c * (p - i * ff);
      sum1 += del1;
      if (fabs(del) < fabs(sum) * kEps) break;
    }
    rkmu = sum;
    rk1 = sum1 * xi2;
    if (scaled) { const double ex = exp(x); rkmu *= ex; rk1 *= ex; }
  } else {
    // Steed's algorithm for the continued fraction CF2
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - xmu2;
    double q = a1, c = a1;
    double a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= 10000; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels / s) < kEps) break;
    }
    h = a1 * h;
    rkmu = sqrt(kPi / (2.0 * x)) / s;
    if (!scaled) rkmu *= exp(-x);
    rk1 = rkmu * (xmu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= t.nl; ++i) {   // K_{mu+i+1} = 2(mu+i)/x K_{mu+i} + K_{mu+i-1}
    const double rktemp = (xmu + i) * xi2 * rk1 + rkmu;
    rkmu = rk1;
    rk1 = rktemp;
  }
  return rkmu;
}

// normcon * s^nu * K_nu(s)   (Matern.cpp:73,80), optionally times exp(s)
__host__ __device__ inline double matern_general_scaled(double s, const CovTable& t, bool scaled) {
  return t.normcon * pow(s, t.nu) * bessel_k_general(s, t, scaled);
}
__device__ inline double matern_general_direct(double s, const CovTable& t) {
  return matern_general_scaled(s, t, false);
}

// value of the tabulated function at squared distance w
__host__ __device__ inline double cov_table_target(double w, double inv_range, double w_split,
                                                   const CovTable& t) {
  const double s = sqrt(w) * inv_range;
  return matern_general_scaled(s, t, w >= w_split);
}

// Lower edge and centre (as mantissas in [1,2)) of interval `idx`; octave exponent in *e.
__host__ __device__ inline void table_interval(int idx, int idx0, double* w_lo, double* w_mid,
                                               double* w_half) {
  const int code = idx + idx0;                        // hi32(w) >> (20 - sub_bits)
  const int hi_lo = code << (20 - kTabSubBits);
  const double lo = from_hilo(hi_lo, 0);
  const double nxt = from_hilo((code + 1) << (20 - kTabSubBits), 0);
  *w_lo = lo;
  *w_half = 0.5 * (nxt - lo);
  *w_mid = lo + *w_half;
}

// Fit one interval: Chebyshev interpolation at kTabDeg+1 first-kind nodes -> monomial
// coefficients in the variable v = mantissa(w) - mantissa(centre).  fvals[i] = f at node i.
__host__ __device__ inline void cheb_fit_to_monomial(const double* fvals, double half_mant,
                                                     double* mono) {
  constexpr int n = kTabDeg + 1;
  const double kPi = 3.141592653589793238462643383279502884;
  double c[n];
  for (int k = 0; k < n; ++k) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += fvals[i] * cos(kPi * k * (i + 0.5) / n);
    c[k] = acc * (2.0 / n);
  }
  c[0] *= 0.5;
  // Chebyshev -> monomial in x in [-1,1]
  double tkm1[n], tk[n], tn[n];
  for (int j = 0; j < n; ++j) { tkm1[j] = 0.0; tk[j] = 0.0; mono[j] = 0.0; }
  tkm1[0] = 1.0;                     // T0
  mono[0] += c[0];
  if (n > 1) {
    tk[1] = 1.0;                     // T1
    mono[1] += c[1];
  }
  for (int k = 2; k < n; ++k) {
    for (int j = 0; j < n; ++j) tn[j] = -tkm1[j];
    for (int j = 0; j + 1 < n; ++j) tn[j + 1] += 2.0 * tk[j];
    for (int j = 0; j < n; ++j) { mono[j] += c[k] * tn[j]; tkm1[j] = tk[j]; tk[j] = tn[j]; }
  }
  // x = v / half_mant  (half_mant is a power of two: exact scaling)
  double sc = 1.0;
  for (int j = 0; j < n; ++j) { mono[j] *= sc; sc /= half_mant; }
}

// One block (32 threads) per interval.  Compiled in gpv_capi.cu only.
#ifdef GPV_DEFINE_TABLE_BUILDER
__global__ void build_cov_table_kernel(CovTable t, double inv_range, double* coef_out) {
  constexpr int n = kTabDeg + 1;
  __shared__ double f[n];
  const int idx = blockIdx.x;
  double w_lo, w_mid, w_half;
  table_interval(idx, t.idx0, &w_lo, &w_mid, &w_half);
  const double kPi = 3.141592653589793238462643383279502884;
  if (threadIdx.x < n) {
    const double xnode = cos(kPi * (threadIdx.x + 0.5) / n);
    const double w = w_mid + w_half * xnode;
    // an interval is "scaled" as a whole: decide by its lower edge (w_split is an interval edge)
    const double s = sqrt(w) * inv_range;
    f[threadIdx.x] = matern_general_scaled(s, t, w_lo >= t.w_split);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mono[n];
    // mantissa half width of every interval: 2^-(sub_bits+1)
    cheb_fit_to_monomial(f, 1.0 / (double)(2 * kTabSub), mono);
    // coefficients are for v in mantissa units: w = 2^e * (mc + v) -> absorb nothing, f is a
    // function of the mantissa within a fixed octave, so no extra scaling is needed.
    for (int k = 0; k < n; ++k) coef_out[tab_coef_index(k, idx)] = mono[k];
  }
}
#endif  // GPV_DEFINE_TABLE_BUILDER

// Fast path: every lane of the warp is strictly inside the table and below the exp split.
__device__ __forceinline__ double cov_general_fast(double r2, int idx, const CovTable& t) {
  const int hi = __double2hiint(r2), lo = __double2loint(r2);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const int keep = 0x000fffff & ~((1 << (20 - kTabSubBits)) - 1);
  const double mc = __hiloint2double((hi & keep) | (1 << (19 - kTabSubBits)) | 0x3ff00000, 0);
  const double v = m - mc;
  const double* cf = t.coef + idx;
  double acc = __ldg(cf + kTabDeg * kTabStride);
#pragma unroll
  for (int k = kTabDeg - 1; k >= 0; --k) acc = fma(acc, v, __ldg(cf + k * kTabStride));
  return acc;
}
// The same polynomial with the coefficients read from the shared-memory window of the table (u_band.cuh).  The
// window is stored as two 32-bit planes (high words, low words), coefficient-major with GW intervals per row, and a
// coefficient is two LDS.32: a warp's lanes sit on a few dozen neighbouring intervals, i.e. on DIFFERENT 4-byte
// banks (32 of them), one wavefront per load.  As 8-byte words the same gather hits the birthday problem of 16 lanes
// on 16 bank pairs: measured 3.0 wavefronts per LDS.64 (profiles/r02_u_band_general_*), 2.4 from global memory.
template <int GW>
__device__ __forceinline__ double cov_general_fast_shared(double r2, const unsigned* __restrict__ hi_w,
                                                          const unsigned* __restrict__ lo_w) {
  const int hi = __double2hiint(r2), lo = __double2loint(r2);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const int keep = 0x000fffff & ~((1 << (20 - kTabSubBits)) - 1);
  const double mc = __hiloint2double((hi & keep) | (1 << (19 - kTabSubBits)) | 0x3ff00000, 0);
  const double v = m - mc;
  auto cf = [&](int k) { return __hiloint2double((int)hi_w[k * GW], (int)lo_w[k * GW]); };
  // even / odd halves in u = v^2: two independent Horner chains instead of one of deg + 1 terms
  constexpr int KE = (kTabDeg / 2) * 2, KO = ((kTabDeg - 1) / 2) * 2 + 1;   // highest even / odd power
  const double u = v * v;
  double ev = cf(KE), od = cf(KO);
#pragma unroll
  for (int k = KE - 2; k >= 0; k -= 2) ev = fma(ev, u, cf(k));
#pragma unroll
  for (int k = KO - 2; k >= 1; k -= 2) od = fma(od, u, cf(k));
  return fma(od, v, ev);
}
// Slow path (per lane correct for anything): zero distance, far pairs (exp split), arguments outside
// the table, NaN.
static __device__ __noinline__ double cov_general_slow(double r2, const UParams& q, const double* __restrict__ etab) {
  const CovTable& t = q.tab;
  if (r2 == 0.0) return q.c0;                                      // Matern.cpp:76-77
  const int idx = (__double2hiint(r2) >> (20 - kTabSubBits)) - t.idx0;
  if ((unsigned)idx >= (unsigned)t.nint) return matern_general_direct(sqrt(r2) * q.inv_range, t);
  double acc = cov_general_fast(r2, idx, t);
  if (r2 >= t.w_split) acc *= exp_neg(sqrt_pos(r2) * q.inv_range, etab);
  return acc;
}
__device__ __forceinline__ bool cov_general_special(double r2, const CovTable& t, int* idx) {
  *idx = (__double2hiint(r2) >> (20 - kTabSubBits)) - t.idx0;
  return ((unsigned)*idx >= (unsigned)t.nint) || (r2 >= t.w_split);   // also catches 0, NaN, Inf
}
__device__ __forceinline__ double cov_general(double r2, const UParams& q, const double* __restrict__ etab) {
  int idx;
  const bool special = cov_general_special(r2, q.tab, &idx);
  if (__any_sync(__activemask(), special)) return cov_general_slow(r2, q, etab);
  return cov_general_fast(r2, idx, q.tab);
}

}  // namespace gpv
