// u_kernels.cuh -- sm_100a kernels for the U_NZentries hot path.
//
// Replaces (reference GPvecchia 0.1.8): the OpenMP loop body src/U_NZentries.cpp:39-69
// (gather -> calcPWD src/dist.cpp:20-30 -> MaternFun src/Matern.cpp:24-86 / EsqeFun
// src/Esqe.cpp:17-39 -> chol(.,"upper") + solve(R, e_last) :60-63) and the numerator of
// vecchia_likelihood_U (R/vecchia_likelihood.R:74-76).
//
// Design (DESIGN.md has the measurements that led here):
//   * a lane group of G lanes owns one conditioning set; P = padded set size (compile time) and
//     G >= ceil(P/2): G = 16 for P <= 32 (two sets per warp), 32 for P <= 64, 8 / 4 for small P.
//   * FOLDED ROWS: lane q keeps two rows of the lower triangle in registers, row q ("low") and row
//     P-1-q ("high"), so every lane carries ~P+1 entries: the triangular load imbalance of a
//     row-per-lane Cholesky disappears and a warp instruction serves 32/G sets.
//   * missing neighbours become LEADING identity padding, so "self is last" (index P-1) holds and
//     the compaction is exactly `inds.elem(find(inds))` (U_NZentries.cpp:44).
//   * the P(P-1)/2 unique covariances are evaluated balanced (point i handles (i, i+t mod P),
//     t = 1..P/2; each lane runs its two points in the same iteration for ILP), staged in shared
//     memory column-major (even stride: conflict-free stores), then pulled into registers.
//   * right-looking Cholesky: column k of L is published to shared memory (contiguous) and read
//     back as group-broadcast 16-byte loads; one DFMA per (k, j) per row kind, no shuffles in the
//     update loop.  x = L^{-T} e_P is a column sweep over the same shared copy (odd stride:
//     conflict-free), which is the reference's solve(R, onevec).
//   * sqrt / exp / rsqrt are branch-free inline sequences (MUFU seed + FMA refinement): the
//     kernel is issue-slot and fp64-pipe bound, so library slow paths are not affordable.
// Tensor cores are not used: batched 31x31 fp64 factorisations, not a dense contraction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include "exp_coeffs.h"

#ifndef GPV_MAX_D
#define GPV_MAX_D 8
#endif

namespace gpv {

enum CovKind : int { COV_EXP = 0, COV_M15 = 1, COV_M25 = 2, COV_ESQE = 3, COV_GENERAL = 4 };

// Piecewise-polynomial table of the general-nu Matern (built per call, bessel_table.cuh).
struct CovTable {
  const double* coef;   // [deg+1][nint], coefficient-major
  int nint;
  int idx0;             // (hi32(w) >> (20 - sub_bits)) - idx0 = interval index
  int sub_bits;
  int deg;
  double w_split;       // w >= w_split : table holds exp(+s) * cov, multiply by exp(-s)
  int win0;             // first interval of the window the band kernels keep in shared memory (u_band.cuh)
  double nu, xmu, gam1, gam2, gampl, gammi, normcon;   // direct (out-of-table) evaluator
  int nl;
};

struct UParams {
  int64_t nrows;          // rows of this shard
  int64_t nsets;          // conditioning sets handed to this launch
  int64_t set_base;       // index of the launch's first set (chunked launches); row = set_base + s if no rowmap
  const int32_t* rowmap;  // [nsets] shard-local row of each set, or nullptr (set s is row s)
  int64_t row0;           // global index of the first row of the shard
  int p;                  // actual set size m+1 (<= P)
  int d;                  // spatial dimension
  const double* locs;     // [Nlocs][d] row-major
  const int32_t* nn;      // [nsets][p] row-major, 0-based ids, -1 = missing
  const uint64_t* cond;   // [nsets] bit j = (revCond[row, j] == TRUE)
  const double* nuggets;  // [Nlocs]
  double* out;            // U values or nullptr
  const int64_t* row_off; // packed offsets [nrows] or nullptr (-> row*p, zero filled)
  const double* zloc;     // [Nlocs] z of each location (0 where unobserved) or nullptr: expanded per call from
                          // zord / obsrank so the set kernel gathers it like a nugget
  int full_z;             // 1: layout is pure `z` conditioning -> also accumulate the denominator terms
  int64_t skip_rows;      // global rows < skip_rows do not enter the likelihood sums
  double* partials;       // [gridDim.x][4] or nullptr: quad.num, sum log x_self, quad.denom, logdet.denom parts
  unsigned long long* nfail;
  long long* first_fail;
  int cov;                // CovKind
  double c0;              // covariance at distance 0
  double c1, c2, c3, c4;  // kind-specific constants (host: setup_cov)
  double inv_range;       // general branch
  CovTable tab;
};

constexpr int kThreadsPerBlock = 128;
constexpr int kWarpsPerBlock = kThreadsPerBlock / 32;
#ifndef GPV_SOLVE_BLOCK
#define GPV_SOLVE_BLOCK 1
#endif
constexpr int kSolveBlock = GPV_SOLVE_BLOCK;   // rows per shuffle round of the back substitution

// --------------------------------------------------------------------------------------------
// branch-free fp64 primitives
// --------------------------------------------------------------------------------------------
// Polynomial / reduction constants live in constant memory so DFMA takes them as c[bank][offset]
// operands; as literals ptxas re-materialised them with two UMOVs each inside the pair loop
// (8% of all issued instructions in the first profile, profiles/r01_*).
__constant__ double kMathC[10] = {
    GPV_EXP_Q0, GPV_EXP_Q1, GPV_EXP_Q2, GPV_EXP_Q3,
    GPV_EXP_64_OVER_LN2,           // [4]
    -GPV_EXP_LN2_64_HI,            // [5]
    -GPV_EXP_LN2_64_LO,            // [6]
    1.0e-300,                      // [7] sqrt guard
    -GPV_EXP_LN2_64,               // [8] one-FMA reduction of the pair stage (exp_negarg_fast_n)
    0.0};
// GPV_PAIR_FAST = 1: the closed-form pair stage uses the shorter sqrt / exp sequences below
// (neg_sqrt_fast_n, exp_negarg_fast_n): 21 instead of 26 fp64 instructions per Matern-1.5 pair.
#ifndef GPV_PAIR_FAST
#define GPV_PAIR_FAST 1
#endif
__constant__ double kExp2Tab[64] = GPV_EXP2_TAB_INIT;   // 2^(j/64), copied to shared memory per block

#ifndef GPV_MINB21
#define GPV_MINB21 4
#endif
#ifndef GPV_MINB26
#define GPV_MINB26 4
#endif
#ifndef GPV_MINB41
#define GPV_MINB41 3
#endif
#ifndef GPV_MINB32
#define GPV_MINB32 4
#endif

__device__ __forceinline__ double rsqrt_seed(double a) {
#ifdef GPV_SIMT_EMU
  return emu_rsqrt_seed(a);                                  // host model of the seed (tests/simt_emu)
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));   // MUFU.RSQ64H, ~20 good bits
  return y;
#endif
}
// 1/sqrt(a) for finite normal a > 0 to ~1 ulp: cubic step (+ a Newton step).
// a <= 0 / NaN / Inf: garbage, callers test `a > 0` themselves (nuggets are clamped at 1e300).
__device__ __forceinline__ double rsqrt_pos(double a) {
  const double y0 = rsqrt_seed(a);
  const double t = a * y0;
  const double e = fma(-t, y0, 1.0);
  double y = fma(fma(e, 0.375, 0.5), e * y0, y0);            // error ~ e^3
  const double t1 = a * y;                                   // once per set: keep the Newton step
  const double e1 = fma(-t1, y, 1.0);
  y = fma(0.5 * e1, y, y);
  return y;
}
// An infinite nugget (Vecchia-Laplace marks missing data with nuggets = Inf,
// R/vecchia_laplace_NR.R:108) must decouple that neighbour (x_j = 0) without producing Inf * 0 in
// the factorisation: 1e300 does the same to 300 digits and keeps the pivot code select-free.
// NaN (Inf * 0 from a latent-conditioned Inf nugget, U_NZentries.cpp:47) passes through and fails
// the row like dpotrf does.
__device__ __forceinline__ double clamp_nugget(double v) { return (v > 1.0e300) ? 1.0e300 : v; }
// 1/a for finite normal a > 0 to ~1 ulp: one cubic step from the MUFU.RCP64H seed (measured seed
// error 2^-19.9, tools/microbench/lat.cu); a <= 0 / NaN / Inf: garbage, callers test the pivot.
__device__ __forceinline__ double rcp_pos(double a) {
  double y0;
#ifdef GPV_SIMT_EMU
  y0 = emu_rcp_seed(a);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
#endif
  const double e = fma(-a, y0, 1.0);
  return fma(fma(e, e, e), y0, y0);                          // y0 (1 + e + e^2): error ~ e^3 = 2^-60
}
// sqrt(w) for w > 0 to ~1 ulp: cubic rsqrt + one Goldschmidt correction, no select.  Callers add
// kSqrtGuard = 1e-300 to the squared distance (fused into its last FMA), which keeps w == 0
// (duplicate locations) away from rsqrt(0) = Inf: sqrt(1e-300) = 1e-150 rounds away in every use
// (exp(-1e-150 c) == 1, 1 + 1e-150 c == 1), so a zero distance still returns exactly c0 like the
// reference's `dist == 0` branches; NaN propagates.
__device__ __forceinline__ double sqrt_pos(double w) {
  const double y0 = rsqrt_seed(w);
  const double t = w * y0;
  const double e = fma(-t, y0, 1.0);
  const double y1 = fma(fma(e, 0.375, 0.5), e * y0, y0);
  const double g = w * y1;
  const double h = 0.5 * y1;
  const double r = fma(-g, h, 0.5);
  return fma(g, r, g);
}
// exp(-s) for s >= 0 (clamped at s = 700: below 1e-304 either way), ~1 ulp, no branches:
// -s = (64 k + j) ln2/64 + r, |r| <= ln2/128; exp(-s) = 2^k * etab[j] * (1 + p(r)), p of degree 5.
// etab = 2^(j/64) in shared memory (a 64-entry gather; constant memory would serialise it).
template <int N>
__device__ __forceinline__ void exp_negarg_n(const double (&ns_in)[N], double (&out)[N], const double* __restrict__ etab);
__device__ __forceinline__ double exp_neg(double s, const double* __restrict__ etab) {
  const double ns[1] = {0.0 - s};   // not -s: a NaN must keep the sign bit clear for the integer clamp
  double e[1];
  exp_negarg_n<1>(ns, e, etab);
  return e[0];
}

// --------------------------------------------------------------------------------------------
// covariance as a function of the SQUARED distance r2 (src/Matern.cpp, src/Esqe.cpp restated;
// every closed form returns exactly c0 at r2 == 0, as the reference's `dist == 0` branches do)
// --------------------------------------------------------------------------------------------
__device__ double matern_general_direct(double s, const CovTable& t);  // bessel_table.cuh
__device__ __forceinline__ double cov_general(double r2, const UParams& q, const double* __restrict__ etab);
__device__ __forceinline__ bool cov_general_special(double r2, const CovTable& t, int* idx);
__device__ __forceinline__ double cov_general_fast(double r2, int idx, const CovTable& t);
static __device__ __noinline__ double cov_general_slow(double r2, const UParams& q, const double* __restrict__ etab);

template <int KIND>
__device__ __forceinline__ double cov_eval(double r2, const UParams& q, const double* __restrict__ etab) {
  if (KIND == COV_EXP) {          // Matern.cpp:38-39   sig2 exp(-d/range)               c1 = 1/range
    return q.c0 * exp_neg(sqrt_pos(r2) * q.c1, etab);
  } else if (KIND == COV_M15) {   // :51-52  sig2 (1+sqrt3 s) exp(-sqrt3 s)              c1 = sqrt3/range
    const double t = sqrt_pos(r2) * q.c1;
    return fma(q.c0, t, q.c0) * exp_neg(t, etab);
  } else if (KIND == COV_M25) {   // :66-68  sig2 exp(-t)(1 + t + t^2/3), t = sqrt5 s     c1 = sqrt5/range
    const double t = sqrt_pos(r2) * q.c1;
    return q.c0 * exp_neg(t, etab) * fma(t, fma(t, 1.0 / 3.0, 1.0), 1.0);
  } else if (KIND == COV_ESQE) {  // Esqe.cpp:32-34  c4 = sig2_1, c1 = 1/r1, c2 = sig2_2, c3 = 1/r2^2
    return fma(q.c2, exp_neg(r2 * q.c3, etab), q.c4 * exp_neg(sqrt_pos(r2) * q.c1, etab));
  } else {
    return cov_general(r2, q, etab);
  }
}

// --------------------------------------------------------------------------------------------
// N covariances at a time, written stage by stage across the N independent chains: the same
// operations in the same order per chain as cov_eval (bit-identical results), but emitted
// interleaved.  ptxas keeps four inlined cov_eval bodies largely one after the other, which leaves
// a warp waiting on its own DFMA latency (profiles: `wait` was the top stall of the quad kernel).
// The argument of the exponential is carried negated (ns = -s): the clamp then selects on what the
// two FMAs of the range reduction consume, and no DADD is spent on a negation.
// --------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void sqrt_pos_n(const double (&w)[N], double (&out)[N]) {
  double y0[N], t[N], e[N], a[N], b2[N], y1[N], g[N], h[N], r[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y0[i] = rsqrt_seed(w[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = w[i] * y0[i];
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-t[i], y0[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = fma(e[i], 0.375, 0.5); b2[i] = e[i] * y0[i]; }
#pragma unroll
  for (int i = 0; i < N; ++i) y1[i] = fma(a[i], b2[i], y0[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) { g[i] = w[i] * y1[i]; h[i] = 0.5 * y1[i]; }
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(-g[i], h[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(g[i], r[i], g[i]);
}
// exp(ns) for ns <= 0 (clamped at -700), N at a time; same arithmetic as exp_neg(-ns)
template <int N>
__device__ __forceinline__ void exp_negarg_n(const double (&ns_in)[N], double (&out)[N],
                                             const double* __restrict__ etab) {
  const double kShift = 6755399441055744.0;
  double ns[N], t[N], kf[N], r[N], T[N], r2[N], qq[N], p[N], v[N];
  int n[N];
  // clamp at -700 on the integer pipe: ns <= -0 or NaN, so the high word read as unsigned grows with
  // |ns|; min(hi, hi(-700)) leaves a (positive, canonical) NaN alone and turns -Inf into -700
#pragma unroll
  for (int i = 0; i < N; ++i)
    ns[i] = __hiloint2double((int)min((unsigned)__double2hiint(ns_in[i]), 0xC085E000u), __double2loint(ns_in[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(ns[i], kMathC[4], kShift);
#pragma unroll
  for (int i = 0; i < N; ++i) { kf[i] = t[i] - kShift; n[i] = __double2loint(t[i]); }
#pragma unroll
  for (int i = 0; i < N; ++i) { r[i] = fma(kf[i], kMathC[5], ns[i]); T[i] = etab[n[i] & 63]; }
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(kf[i], kMathC[6], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) { r2[i] = r[i] * r[i]; qq[i] = fma(kMathC[3], r[i], kMathC[2]); }
#pragma unroll
  for (int i = 0; i < N; ++i) qq[i] = fma(qq[i], r[i], kMathC[1]);
#pragma unroll
  for (int i = 0; i < N; ++i) qq[i] = fma(qq[i], r[i], kMathC[0]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(qq[i], r2[i], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = fma(T[i], p[i], T[i]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    out[i] = __hiloint2double(__double2hiint(v[i]) + (n[i] >> 6) * 1048576, __double2loint(v[i]));
}
// The pair stage's own sqrt and exp (GPV_PAIR_FAST).  The kernel is bound by the fp64 pipe and 56 % of
// its fp64 instructions are these two functions (profiles/r01_u_band_closed_*), so they are cut to what
// the parity bar needs -- about one ulp of the ARGUMENT, which is what the reference's own
// sqrt -> divide -> multiply chain carries (Matern.cpp:38-68):
//  * -sqrt(w) = g (1 + e/2 + 3 e^2/8), g = -w y0, e = 1 - w y0^2 from the MUFU.RSQ64H seed y0
//    (|e| < 2^-18.9: the dropped term 5 e^3/16 is below 2^-58).  The rounding error d of g enters e
//    with the opposite sign, so half of it cancels: 5 fp64 instructions instead of the 9 of
//    sqrt_pos_n, and a dependency chain of 4 instead of 7.
//  * exp: r = ns - kf c in ONE fma with c = nearest(ln2/64); the product is exact inside the fma, what is
//    lost is kf (ln2/64 - c) <= |ns| 2^-53, i.e. half an ulp of the argument (absolute error of the
//    result <= s e^-s 2^-53 <= 0.37 ulp(1)).  2 instead of 3 instructions; polynomial as before.
// Verified on the host against mpmath over the seed-error envelope: tools/check_pair_fast.py.
template <int N>
__device__ __forceinline__ void neg_sqrt_fast_n(const double (&w)[N], double (&out)[N]) {
  double y0[N], g[N], e[N], a[N], ge[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y0[i] = rsqrt_seed(w[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = w[i] * (-y0[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(g[i], y0[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = fma(e[i], 0.375, 0.5); ge[i] = g[i] * e[i]; }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(a[i], ge[i], g[i]);
}
// ES: stride of the 2^(j/64) table.  ES = 1: 64 doubles; ES = 16: every entry replicated 16 times and `etab`
// already offset by lane % 16, so the 16 lanes of a half-warp read 16 different bank pairs whatever their j
// (profiles/r02_*: the 64-entry table cost 4.8 wavefronts per lookup instead of 2).
template <int N, int ES = 1>
__device__ __forceinline__ void exp_negarg_fast_n(const double (&ns_in)[N], double (&out)[N],
                                                  const double* __restrict__ etab) {
  const double kShift = 6755399441055744.0;
  double ns[N], t[N], kf[N], r[N], T[N], r2[N], qq[N], p[N], v[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i)   // clamp at -700 on the integer pipe, see exp_negarg_n
    ns[i] = __hiloint2double((int)min((unsigned)__double2hiint(ns_in[i]), 0xC085E000u), __double2loint(ns_in[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(ns[i], kMathC[4], kShift);
#pragma unroll
  for (int i = 0; i < N; ++i) { kf[i] = t[i] - kShift; n[i] = __double2loint(t[i]); }
#pragma unroll
  for (int i = 0; i < N; ++i) { r[i] = fma(kf[i], kMathC[8], ns[i]); T[i] = etab[(n[i] & 63) * ES]; }
#pragma unroll
  for (int i = 0; i < N; ++i) { r2[i] = r[i] * r[i]; qq[i] = fma(kMathC[3], r[i], kMathC[2]); }
#pragma unroll
  for (int i = 0; i < N; ++i) qq[i] = fma(qq[i], r[i], kMathC[1]);
#pragma unroll
  for (int i = 0; i < N; ++i) qq[i] = fma(qq[i], r[i], kMathC[0]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(qq[i], r2[i], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = fma(T[i], p[i], T[i]);
#pragma unroll
  for (int i = 0; i < N; ++i)
    out[i] = __hiloint2double(__double2hiint(v[i]) + (n[i] >> 6) * 1048576, __double2loint(v[i]));
}
struct CovConsts { double c0, c1, c2, c3, c4; };   // what the closed forms read of UParams, by value
template <int KIND, int N, class C, int ES = 1>
__device__ __forceinline__ void cov_eval_n(const double (&r2)[N], double (&v)[N], const C& q,
                                           const double* __restrict__ etab) {
  static_assert(KIND != COV_GENERAL, "closed forms only");
  static_assert(GPV_PAIR_FAST || ES == 1, "the replicated exp table is read by exp_negarg_fast_n only");
  double sq[N], ns[N], e[N];
#if GPV_PAIR_FAST
  neg_sqrt_fast_n<N>(r2, sq);
#pragma unroll
  for (int i = 0; i < N; ++i) ns[i] = sq[i] * q.c1;
  exp_negarg_fast_n<N, ES>(ns, e, etab);
#else
  sqrt_pos_n<N>(r2, sq);
#pragma unroll
  for (int i = 0; i < N; ++i) ns[i] = sq[i] * (-q.c1);
  exp_negarg_n<N>(ns, e, etab);
#endif
  if (KIND == COV_EXP) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = q.c0 * e[i];
  } else if (KIND == COV_M15) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fma(-q.c0, ns[i], q.c0) * e[i];
  } else if (KIND == COV_M25) {
#pragma unroll
#if GPV_PAIR_FAST
    for (int i = 0; i < N; ++i) v[i] = e[i] * fma(ns[i], fma(ns[i], q.c0 * (1.0 / 3.0), -q.c0), q.c0);   // c0/3: loop invariant
#else
    for (int i = 0; i < N; ++i) v[i] = q.c0 * e[i] * fma(ns[i], fma(ns[i], 1.0 / 3.0, -1.0), 1.0);
#endif
  } else {
    double ns2[N], e2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) ns2[i] = r2[i] * (-q.c3);
#if GPV_PAIR_FAST
    exp_negarg_fast_n<N, ES>(ns2, e2, etab);
#else
    exp_negarg_n<N>(ns2, e2, etab);
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fma(q.c2, e2[i], q.c4 * e[i]);
  }
}

// Packed lower triangle, column-major: column k holds rows k..P-1 at tri_col(k) + (r - k).  Used both
// for the staged covariance matrix and, column by column, for L (a column of L overwrites the
// staged column it was computed from, which is dead by then).  496 doubles for P = 31 instead of
// the 992 of a square buffer: shared memory no longer caps residency below the register limit.
__host__ __device__ constexpr int tri_col(int k, int P) { return k * P - (k * (k - 1)) / 2; }

template <int G, int P, int D>
struct SetLayout {
  static constexpr int DD = (D > 0) ? D : GPV_MAX_D;
  static constexpr int NLOW = (P + 1) / 2;                 // rows 0..NLOW-1, row q on lane q
  static constexpr int NHIGH = P / 2;                      // rows P-1..NLOW, row P-1-q on lane q
  static constexpr int kScratch = tri_col(P, P);           // per-lane dump slots for inactive pair stores
  static constexpr int kBuf = ((tri_col(P, P) + G + 1) / 2) * 2;
  static constexpr int kT = P / 2;                         // pair-stage iterations
  static constexpr int PX = ((P + 1) / 2) * 2;             // coordinate row stride
  static constexpr int kX = DD * PX;                       // coordinates of the P points
  static constexpr int kNug = PX;                          // nuggets of the P points
  static constexpr int kZ = PX;                            // z of the P points (likelihood only)
  static constexpr int kI = ((PX / 2 + 1) / 2) * 2;        // P int32 compacted ids (kept even)
  static constexpr int kRawI = G;                          // 2G int32 raw ids of a row (as stored)
  static constexpr int kMeta = 2;                          // cond mask (8 B), row (4 B), pad
  // one input stage = everything the pair stage and the epilogue read about a set; two of them so
  // that cp.async fills stage b^1 while stage b is being factored
  static constexpr int kStage = kX + kNug + kZ + kI + kRawI + kMeta;
  static constexpr int kOffNug = kX, kOffZ = kX + kNug, kOffIds = kX + kNug + kZ,
                       kOffRaw = kX + kNug + kZ + kI, kOffMeta = kX + kNug + kZ + kI + kRawI;
  static constexpr int kSetsPerWarp = 32 / G;
  // offset consecutive sets of a warp by 128/kSetsPerWarp bytes (mod 128) so that the per-set
  // broadcast loads of one warp instruction fall into different banks
  static constexpr int kRaw = kBuf + 2 * kStage;
  static constexpr int kWant = (kSetsPerWarp > 1) ? 16 / kSetsPerWarp : 0;   // doubles, mod 16
  static constexpr int kPad = (kSetsPerWarp > 1) ? ((kWant - (kRaw % 16)) + 16) % 16 : (kRaw % 2);
  static constexpr int kDoubles = kRaw + kPad;
  static constexpr int kBytesPerBlock = kDoubles * 8 * kSetsPerWarp * kWarpsPerBlock;
  static_assert(NLOW <= G, "a lane group must hold ceil(P/2) low rows");
  static_assert(P >= 2 && P <= 64 && G <= 32, "unsupported set size");
};

// squared distance between my point (registers) and staged point j, plus `guard` (1e-300 for the
// closed forms, see sqrt_pos; 0 for the general branch, which tests r2 == 0 itself).  D == 2 stages
// the coordinates as double2 (one 16-byte load per partner); other D use one array per coordinate.
template <int D>
__device__ __forceinline__ double pair_r2(const double* __restrict__ xs, int PX, const double* xi, int j,
                                          int d, double guard) {
  double r2 = guard;
  if (D == 2) {
    const double2 v = reinterpret_cast<const double2*>(xs)[j];
    const double dx = xi[0] - v.x, dy = xi[1] - v.y;
    r2 = fma(dy, dy, fma(dx, dx, guard));
  } else if (D > 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double dd = xi[c] - xs[c * PX + j];
      r2 = fma(dd, dd, r2);
    }
  } else {
#pragma unroll
    for (int c = 0; c < GPV_MAX_D; ++c) {
      if (c < d) {
        const double dd = xi[c] - xs[c * PX + j];
        r2 = fma(dd, dd, r2);
      }
    }
  }
  return r2;
}

// Where pair (i, i+t mod P) goes in the packed staged triangle depends only on (i, t): a per-block
// table (built once per launch) holds, for iteration t and lane q, the two store offsets of the
// lane's low and high point as 16-bit halves; inactive combinations point at the dump slot.
template <int G, int P, int D>
__device__ __forceinline__ void build_store_table(unsigned* __restrict__ stab) {
  using LY = SetLayout<G, P, D>;
  for (int idx = threadIdx.x; idx < LY::kT * G; idx += blockDim.x) {
    const int t = idx / G + 1, q = idx % G;
    const bool full = (2 * t < P);                       // even P: t = P/2 is covered by i < P/2 only
    unsigned lo16 = LY::kScratch + q, hi16 = LY::kScratch + q;   // own dump slot: no write-write race
    if (q < LY::NLOW && (full || q < P / 2)) {
      const int i = q;
      int j = i + t; if (j >= P) j -= P;
      const int a = i > j ? i : j, b = i > j ? j : i;
      lo16 = tri_col(b, P) + a - b;
    }
    if (q < LY::NHIGH && (full || (P - 1 - q) < P / 2)) {
      const int i = P - 1 - q;
      int j = i + t; if (j >= P) j -= P;
      const int a = i > j ? i : j, b = i > j ? j : i;
      hi16 = tri_col(b, P) + a - b;
    }
    stab[idx] = lo16 | (hi16 << 16);
  }
}

template <int KIND, int G, int P, int D>
__device__ __forceinline__ void pair_eval_store(const UParams& q, double* __restrict__ As,
                                                const double* __restrict__ xs, const double* xl,
                                                const double* xh, int il, int ih,
                                                const unsigned* __restrict__ stab_lane,
                                                const double* __restrict__ etab, int t, int d) {
  using LY = SetLayout<G, P, D>;
  int jl = il + t; if (jl >= P) jl -= P;
  int jh = ih + t; if (jh >= P) jh -= P;
  const unsigned offs = stab_lane[(t - 1) * G];
  const double guard = (KIND == COV_GENERAL) ? 0.0 : kMathC[7];
  const double r2l = pair_r2<D>(xs, LY::PX, xl, jl, d, guard);
  const double r2h = pair_r2<D>(xs, LY::PX, xh, jh, d, guard);
  double vl, vh;
  if (KIND == COV_GENERAL) {
    // one warp-uniform decision for both evaluations of the iteration (the loop is convergent)
    int idxl, idxh;
    const bool sp = cov_general_special(r2l, q.tab, &idxl) | cov_general_special(r2h, q.tab, &idxh);
    if (__any_sync(0xffffffffu, sp)) {
      vl = cov_general_slow(r2l, q, etab);
      vh = cov_general_slow(r2h, q, etab);
    } else {
      vl = cov_general_fast(r2l, idxl, q.tab);
      vh = cov_general_fast(r2h, idxh, q.tab);
    }
  } else {
    const double r2v[2] = {r2l, r2h};
    double vv[2];
    cov_eval_n<(KIND == COV_GENERAL) ? COV_EXP : KIND, 2>(r2v, vv, q, etab);
    vl = vv[0]; vh = vv[1];
  }
  As[offs & 0xffffu] = vl;
  As[offs >> 16] = vh;
}

// pair stage: point i evaluates the covariances (i, i+t mod P), t = 1..P/2; every lane carries its
// low point q and its high point P-1-q through the same iteration (independent chains for ILP).
// Entries that involve padding are computed on dummy coordinates here and zeroed afterwards:
// only the first m rows of a data set have padding.
template <int KIND, int G, int P, int D>
__device__ __forceinline__ void pair_stage(const UParams& q, double* __restrict__ As,
                                           const double* __restrict__ xs, const double* xl,
                                           const double* xh, int gl, const unsigned* __restrict__ stab,
                                           const double* __restrict__ etab, int d) {
  using LY = SetLayout<G, P, D>;
  const int il = (gl < LY::NLOW) ? gl : 0;
  const int ih = (gl < LY::NHIGH) ? (P - 1 - gl) : (P - 1);
  const unsigned* stab_lane = stab + gl;
  constexpr int T = LY::kT;
  // kept rolled: two evaluations per lane already give two independent chains, and unrolling by two
  // measured within 1%
#pragma unroll 1
  for (int t = 1; t <= T; ++t) pair_eval_store<KIND, G, P, D>(q, As, xs, xl, xh, il, ih, stab_lane, etab, t, d);
}

// resident blocks per SM the register allocation is sized for (shared memory allows the same)
template <int P>
struct Occupancy {
  static constexpr int kMinBlocks = (P <= 21) ? GPV_MINB21 : (P <= 26) ? GPV_MINB26 : (P <= 32) ? GPV_MINB32
                                    : (P <= 41) ? GPV_MINB41 : 1;
};

// GENERAL = false: the closed forms (exp, Matern 1.5 / 2.5, esqe) selected at run time by q.cov;
// GENERAL = true : the general-nu table path.  Separate instantiations so that the general path's
// fallback code (pow / log / sinh / continued fraction) does not weigh on the register allocation of
// the closed-form kernel (measured: 6% on the nu = 1.5 kernel).
template <int G, int P, int D, bool GENERAL>
__global__ void __launch_bounds__(kThreadsPerBlock, Occupancy<P>::kMinBlocks)
u_sets_kernel(const UParams q) {
  using LY = SetLayout<G, P, D>;
  constexpr int NLOW = LY::NLOW, NHIGH = LY::NHIGH, PX = LY::PX;
  constexpr int SETS = LY::kSetsPerWarp;
  constexpr unsigned FULL = 0xffffffffu;

#ifdef GPV_SIMT_EMU
  double* smem = emu_dynamic_smem();                          // provided by the host harness (tests/simt_emu)
#else
  extern __shared__ __align__(16) double smem[];
#endif
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / G;
  const int gl = lane % G;
  const int base = sub * G;
  const unsigned gmask = (G == 32) ? FULL : ((1u << G) - 1u);
  const int d = (D > 0) ? D : q.d;
  const int p = q.p;

  double* buf = smem + (size_t)(warp * SETS + sub) * LY::kDoubles;   // staged triangle, then L
  double* stage0 = buf + LY::kBuf;                                   // two input stages

  __shared__ unsigned stab[(LY::kT > 0 ? LY::kT : 1) * G];
  __shared__ double etab[64];
  build_store_table<G, P, D>(stab);
  if (threadIdx.x < 64) etab[threadIdx.x] = kExp2Tab[threadIdx.x];
  __syncthreads();

  const bool lowv = gl < NLOW;
  const bool highv = gl < NHIGH;
  const int rl = lowv ? gl : NLOW - 1;               // clamped row indices keep idle lanes in bounds
  const int rh = highv ? (P - 1 - gl) : (P - 1);

  double acc_quad = 0.0, acc_logd = 0.0, acc_qden = 0.0, acc_lden = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock * SETS;
  const int64_t first = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * SETS;

  // ---- input pipeline (cp.async, no registers held across the factorisation) ----------------------
  // fetch_raw(s, st): the row of neighbour ids, its revCond mask and its row number -> stage st
  auto fetch_raw = [&](int64_t sidx, double* st) {
    int* raw = reinterpret_cast<int*>(st + LY::kOffRaw);
    double* meta = st + LY::kOffMeta;
    int* metai = reinterpret_cast<int*>(meta + 1);
    if (sidx < q.nsets) {
      const int32_t* nnr = q.nn + sidx * (int64_t)p;
      if (gl < p) __pipeline_memcpy_async(raw + gl, nnr + gl, 4); else raw[gl] = -1;
      if (gl + G < p) __pipeline_memcpy_async(raw + gl + G, nnr + gl + G, 4); else raw[gl + G] = -1;
      if (gl == 0) {
        __pipeline_memcpy_async(meta, q.cond + sidx, 8);
        if (q.rowmap != nullptr) __pipeline_memcpy_async(metai, q.rowmap + sidx, 4);
        else metai[0] = (int)(q.set_base + sidx);
      }
    } else {
      raw[gl] = -1;
      raw[gl + G] = -1;
      if (gl == 0) { reinterpret_cast<unsigned long long*>(meta)[0] = 0ull; metai[0] = -1; }
    }
  };
  // gather(st): compaction of the raw ids (U_NZentries.cpp:41-45: entries gl and gl + G of the row),
  // then coordinates and nuggets of my two points -> stage st.  Returns n0.
  auto gather = [&](double* st) -> int {
    const int* raw = reinterpret_cast<const int*>(st + LY::kOffRaw);
    int* ids = reinterpret_cast<int*>(st + LY::kOffIds);
    double* xs = st;
    double* nug = st + LY::kOffNug;
    const int raw0 = raw[gl], raw1 = raw[gl + G];
    const unsigned b0 = (__ballot_sync(FULL, raw0 >= 0) >> base) & gmask;
    const unsigned b1 = (__ballot_sync(FULL, raw1 >= 0) >> base) & gmask;
    const int n0 = __popc(b0) + __popc(b1);
    const int npad = P - n0;
    const unsigned below = (1u << gl) - 1u;
    if (gl < npad) ids[gl] = -1;
    if (gl + G < npad) ids[gl + G] = -1;
    __syncwarp();
    if (raw0 >= 0) ids[npad + __popc(b0 & below)] = raw0;
    if (raw1 >= 0) ids[npad + __popc(b0) + __popc(b1 & below)] = raw1;
    __syncwarp();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const bool valid = which ? highv : lowv;
      const int r = which ? rh : rl;
      if (!valid) continue;
      const int id = ids[r];
      if (id >= 0) {
        if (D == 2) {
          __pipeline_memcpy_async(xs + 2 * r, q.locs + 2 * (int64_t)id, 16);
        } else {
          for (int c = 0; c < d; ++c) __pipeline_memcpy_async(xs + c * PX + r, q.locs + (int64_t)id * d + c, 8);
        }
        __pipeline_memcpy_async(nug + r, q.nuggets + id, 8);
        if (q.zloc != nullptr) __pipeline_memcpy_async(st + LY::kOffZ + r, q.zloc + id, 8);
      } else {
        if (D == 2) {
          reinterpret_cast<double2*>(xs)[r] = make_double2(0.0, 0.0);
        } else {
          for (int c = 0; c < d; ++c) xs[c * PX + r] = 0.0;
        }
        nug[r] = 0.0;
      }
    }
    return n0;
  };

  int bsel = 0;
  fetch_raw(first + sub, stage0);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncwarp();
  int n0 = gather(stage0);
  fetch_raw(first + stride + sub, stage0 + LY::kStage);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncwarp();

  for (int64_t s0 = first; s0 < q.nsets; s0 += stride) {
    double* st = stage0 + bsel * LY::kStage;
    double* stn = stage0 + (bsel ^ 1) * LY::kStage;
    // meta of the current set into registers, then its raw/meta slots are free for set i+2
    const uint64_t cmask = reinterpret_cast<const unsigned long long*>(st + LY::kOffMeta)[0];
    const int row = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[0];
    const bool row_ok = row >= 0;
    const int npad = P - n0;
    const int n0_next = gather(stn);                       // set i+1: ids now, coordinates in flight
    __syncwarp();
    fetch_raw(s0 + 2 * stride + sub, st);                  // set i+2: ids in flight
    __pipeline_commit();

    // ---- 1./2. my two points of the current set (staged by the previous iteration) ---------------
    const double* xs = st;
    const int* ids = reinterpret_cast<const int*>(st + LY::kOffIds);
    const double* nugs = st + LY::kOffNug;
    double xl[LY::DD], xh[LY::DD];
    double dgl = 1.0, dgh = 1.0;
    int idl = -1, idh = -1;
    bool cl = false, ch = false;
    if (lowv) idl = ids[rl];
    if (highv) idh = ids[rh];
    if (D == 2) {
      const double2 vl = reinterpret_cast<const double2*>(xs)[rl];
      const double2 vh = reinterpret_cast<const double2*>(xs)[rh];
      xl[0] = vl.x; xl[1] = vl.y; xh[0] = vh.x; xh[1] = vh.y;
    } else {
#pragma unroll
      for (int c = 0; c < LY::DD; ++c) {
        xl[c] = (c < d) ? xs[c * PX + rl] : 0.0;
        xh[c] = (c < d) ? xs[c * PX + rh] : 0.0;
      }
    }
    if (idl >= 0) {
      // compacted entry j reads revCond[row, p - n0 + j] (:47); local index = npad + j
      cl = (cmask >> ((rl - (P - p)) & 63)) & 1ull;
      dgl = q.c0 + clamp_nugget(nugs[rl] * (1.0 - (cl ? 1.0 : 0.0)));   // Inf * 0 = NaN kept
    }
    if (idh >= 0) {
      ch = (cmask >> ((rh - (P - p)) & 63)) & 1ull;
      dgh = q.c0 + clamp_nugget(nugs[rh] * (1.0 - (ch ? 1.0 : 0.0)));
    }

    // ---- 3. covariance pairs -> shared staging (column-major lower, even stride) -----------------
    if (GENERAL) {
      pair_stage<COV_GENERAL, G, P, D>(q, buf, xs, xl, xh, gl, stab, etab, d);
    } else {
      switch (q.cov) {
        case COV_EXP: pair_stage<COV_EXP, G, P, D>(q, buf, xs, xl, xh, gl, stab, etab, d); break;
        case COV_M15: pair_stage<COV_M15, G, P, D>(q, buf, xs, xl, xh, gl, stab, etab, d); break;
        case COV_M25: pair_stage<COV_M25, G, P, D>(q, buf, xs, xl, xh, gl, stab, etab, d); break;
        default: pair_stage<COV_ESQE, G, P, D>(q, buf, xs, xl, xh, gl, stab, etab, d); break;
      }
    }
    if (__any_sync(FULL, npad > 0)) {
      // padding occupies the leading indices: every pair with a padded point has its smaller index
      // < npad, i.e. lives in columns 0..npad-1 of the staged lower triangle -> zero those columns
      __syncwarp();
      for (int j = 0; j < npad; ++j) {
        if (lowv && rl > j) buf[tri_col(j, P) + rl - j] = 0.0;
        if (highv && rh > j) buf[tri_col(j, P) + rh - j] = 0.0;
      }
    }
    if (lowv) buf[tri_col(rl, P)] = dgl;
    if (highv) buf[tri_col(rh, P)] = dgh;
    __syncwarp();

    // ---- 4. my two rows of the lower triangle into registers ---------------------------------------
    double lo[NLOW], hi[P];
#pragma unroll
    for (int j = 0; j < NLOW; ++j) lo[j] = buf[tri_col(j, P) + rl - j];   // j > rl: stale, never used
#pragma unroll
    for (int j = 0; j < P; ++j) hi[j] = buf[tri_col(j, P) + rh - j];      // j > rh: stale, never used
    __syncwarp();

    // ---- 5. right-looking LDL^T (square-root-free Cholesky; chol(covmat,"upper"), U_NZentries.cpp:61)
    // Sigma = L D L^T with unit lower L.  Column k is published to shared memory already divided by
    // the pivot (w[j] = a[j][k] / d_k = L[j][k]); every lane keeps its own UNSCALED entry c = a[r][k]
    // and updates a[r][j] -= c * w[j].  The reference's factor is R = D^{1/2} L^T, so its
    // R^{-1} e_P = L^{-T} e_P / sqrt(d_P): one rsqrt per set instead of one per pivot, and a
    // unit-triangular column sweep.  A pivot that is not > 0 (or NaN) is dpotrf's failure.
    bool fail = false;
    double dlast = 1.0;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double akk = (k < NLOW) ? __shfl_sync(FULL, lo[k < NLOW ? k : 0], base + k)
                                    : __shfl_sync(FULL, hi[k], base + (P - 1 - k));
      // positive, normal, finite -- dpotrf's `ajj <= 0 || isnan(ajj)` test on the integer pipe
      fail = fail || ((unsigned)(__double2hiint(akk) - 0x00100000) >= 0x7fe00000u);
      if (k == P - 1) { dlast = akk; break; }
      const double inv = rcp_pos(akk);      // an Inf nugget arrives here as 1e300 (clamp_nugget)
      // publish L[r][k] = a[r][k] / d_k for r >= k (a packed column must not be written above its top)
      const int ck = tri_col(k, P) - k;     // L[r][k] at buf[ck + r]
      double cl = 0.0;
      if (k < NLOW) {
        cl = lo[k < NLOW ? k : 0];          // lanes gl >= k: a[gl][k]
        if (gl >= k && lowv) buf[ck + rl] = cl * inv;
      }
      const double chh = hi[k];             // lanes with rh >= k: a[rh][k]
      if (gl <= P - 1 - k && highv) buf[ck + rh] = chh * inv;
      __syncwarp();
      // trailing update: a[r][j] -= a[r][k] * L[j][k], j = k+1..P-1 (garbage beyond the row end)
      int j = k + 1;
      if (j < P && ((ck + j) & 1) != 0) {   // odd offset: one scalar broadcast load first
        const double l1 = buf[ck + j];
        if (j < NLOW) lo[j < NLOW ? j : 0] = fma(-cl, l1, lo[j < NLOW ? j : 0]);
        hi[j] = fma(-chh, l1, hi[j]);
        ++j;
      }
#pragma unroll
      for (; j + 1 < P; j += 2) {           // 16-byte aligned broadcast loads for (j, j+1)
        const double2 l2 = *reinterpret_cast<const double2*>(&buf[ck + j]);
        if (j < NLOW) lo[j < NLOW ? j : 0] = fma(-cl, l2.x, lo[j < NLOW ? j : 0]);
        if (j + 1 < NLOW) lo[j + 1 < NLOW ? j + 1 : 0] = fma(-cl, l2.y, lo[j + 1 < NLOW ? j + 1 : 0]);
        hi[j] = fma(-chh, l2.x, hi[j]);
        hi[j + 1] = fma(-chh, l2.y, hi[j + 1]);
      }
      if (j < P) {
        const double l1 = buf[ck + j];
        if (j < NLOW) lo[j < NLOW ? j : 0] = fma(-cl, l1, lo[j < NLOW ? j : 0]);
        hi[j] = fma(-chh, l1, hi[j]);
      }
    }

    // ---- 6. x = L^{-T} e_P / sqrt(d_P)  (solve(R, onevec), U_NZentries.cpp:62): unit-triangular column
    // sweep, j = P-1..1, on t = -y: t_r = -sum_{j > r} L[j][r] t_j, and the unit right-hand side enters as
    // t_{P-1} = -1 (row P-1 is lane 0's high row and is never updated below).  The sweep is a chain of
    // shuffle -> FMA latencies, so it runs kSolveBlock rows per round: the partial sums of the block's
    // rows are shuffled together, every lane finishes the block's small triangle redundantly from
    // broadcast loads, and then applies the block to its own two rows.  Same operations in the same order
    // as the row-at-a-time sweep, a quarter of the dependent shuffles.
    double sl = 0.0, sh = (gl == 0) ? -1.0 : 0.0;
    const int cbl = tri_col(rl, P) - rl, cbh = tri_col(rh, P) - rh;
#pragma unroll
    for (int j = P - 1; j >= 1; j -= kSolveBlock) {
      double t[kSolveBlock];
#pragma unroll
      for (int b = 0; b < kSolveBlock; ++b) {
        const int r = j - b;
        if (r < 1) continue;
        t[b] = (r >= NLOW) ? __shfl_sync(FULL, sh, base + (P - 1 - r)) : __shfl_sync(FULL, sl, base + r);
      }
#pragma unroll
      for (int b = 1; b < kSolveBlock; ++b) {
        const int r = j - b;
        if (r < 1) continue;
#pragma unroll
        for (int a = 0; a < b; ++a)            // L[j-a][r] sits at buf[tri_col(r) - r + (j - a)]: broadcast
          t[b] = fma(-buf[tri_col(r, P) - r + (j - a)], t[a], t[b]);
      }
#pragma unroll
      for (int b = 0; b < kSolveBlock; ++b) {
        const int r = j - b;
        if (r < 1) continue;
        if (rl < r) sl = fma(-buf[cbl + r], t[b], sl);
        if (r > NLOW && rh < r) sh = fma(-buf[cbh + r], t[b], sh);
      }
    }
    const double rs = rsqrt_pos(dlast);
    double xlow = -sl * rs;
    double xhigh = -sh * rs;
    if (fail) { xlow = 0.0; xhigh = 0.0; }   // row stays zero (:64-66)

    // ---- 7. outputs ----------------------------------------------------------------------------------
    if (fail && row_ok && gl == 0 && n0 > 0) {
      atomicAdd(q.nfail, 1ull);
      atomicMin(q.first_fail, (long long)(q.row0 + row));
    }
    if (q.out != nullptr && row_ok) {
      if (q.row_off != nullptr) {
        double* o = q.out + q.row_off[row];
        if (idl >= 0) o[rl - npad] = xlow;
        if (idh >= 0) o[rh - npad] = xhigh;
      } else {
        double* o = q.out + (int64_t)row * p;
        if (idl >= 0) o[rl - npad] = xlow;
        if (idh >= 0) o[rh - npad] = xhigh;
        if (gl >= n0 && gl < p) o[gl] = 0.0;             // zero fill beyond n0 (:33)
        if (gl + G >= n0 && gl + G < p) o[gl + G] = 0.0;
      }
    }
    if (q.partials != nullptr) {
      // quadform.num: (sum_{j: revCond = 0} x_j z_j)^2 ; logdet.num: log x_self (vecchia_likelihood.R:74-76)
      const double* zst = st + LY::kOffZ;
      double t = 0.0;
      if (idl >= 0 && !cl) t = xlow * zst[rl];
      if (idh >= 0 && !ch) t = fma(xhigh, zst[rh], t);
#pragma unroll
      for (int o = G / 2; o >= 1; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
      const double xself = __shfl_sync(FULL, xhigh, base);     // row P-1 is lane 0's high row
      if (gl == 0 && row_ok && n0 > 0 && (q.row0 + row) >= q.skip_rows) {
        acc_quad += t * t;
        acc_logd += log(xself);
        if (q.full_z) {
          // pure `z` conditioning: U_y U_y^T is diagonal, W_kk = x_kk^2 + 1/tau_k, and
          // z2_k = x_kk q_k - z_k / tau_k (vecchia_likelihood.R:85-91 per row)
          const double tau = nugs[P - 1], zk = zst[P - 1];
          const double w = fma(xself, xself, 1.0 / tau);
          const double z2 = fma(xself, t, -zk / tau);
          acc_qden += z2 * z2 / w;
          acc_lden += log(w);
        }
      }
    }
    __pipeline_wait_prior(0);                              // set i+1 staged, ids of set i+2 landed
    __syncwarp();
    n0 = n0_next;
    bsel ^= 1;
  }

  // ---- deterministic block reduction of the likelihood partial sums ---------------------------------
  if (q.partials != nullptr) {
    __shared__ double red[kWarpsPerBlock][4];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      acc_quad += __shfl_xor_sync(FULL, acc_quad, o);
      acc_logd += __shfl_xor_sync(FULL, acc_logd, o);
      acc_qden += __shfl_xor_sync(FULL, acc_qden, o);
      acc_lden += __shfl_xor_sync(FULL, acc_lden, o);
    }
    if (lane == 0) { red[warp][0] = acc_quad; red[warp][1] = acc_logd; red[warp][2] = acc_qden; red[warp][3] = acc_lden; }
    __syncthreads();
    if (threadIdx.x < 4) {
      double a0 = 0.0;
      for (int w = 0; w < kWarpsPerBlock; ++w) a0 += red[w][threadIdx.x];
      q.partials[4 * blockIdx.x + threadIdx.x] = a0;
    }
  }
}

}  // namespace gpv
