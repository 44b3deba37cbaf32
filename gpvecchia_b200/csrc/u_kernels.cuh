// u_kernels.cuh -- sm_100a kernels for the U_NZentries hot path.
//
// Replaces (reference GPvecchia 0.1.8): the OpenMP loop body src/U_NZentries.cpp:39-69
// (gather -> calcPWD src/dist.cpp:20-30 -> MaternFun src/Matern.cpp:24-86 / EsqeFun
// src/Esqe.cpp:17-39 -> chol(.,"upper") + solve(R, e_last) :60-63) and the numerator of
// vecchia_likelihood_U (R/vecchia_likelihood.R:74-76).
//
// Design: one lane-group of G lanes (G = 8/16/32 -> 4/2/1 sets per warp) per
// conditioning set; P = padded set size (compile time, P <= G).
//   1. ids of the row loaded coalesced and compacted like `inds.elem(find(inds))`
//      (U_NZentries.cpp:44); missing entries become LEADING identity padding, so "self is last"
//      (index P-1) is preserved;
//   2. coordinates + nuggets gathered once per neighbour, staged in shared memory;
//   3. the P(P-1)/2 unique covariances are evaluated balanced over the lanes (lane i handles the
//      pairs (i, i+t mod P), t = 1..P/2), staged in shared memory, then every lane pulls ITS ROW
//      of the lower triangle into registers;
//   4. right-looking Cholesky, row-per-lane in registers; column k of L is published through
//      shared memory and consumed by warp-broadcast (vectorised) loads: one DFMA per (k, j) per
//      group, no shuffles in the update loop;
//   5. x = L^{-T} e_P by a column sweep over the shared-memory copy of L (odd leading dimension,
//      conflict-free), which is the reference's solve(R, onevec);
//   6. outputs: U values (row-major zero-filled, or packed createU.R:158-160 order) and/or the
//      fused likelihood terms, reduced deterministically per block.
// Bound by the fp64 FMA pipe; tensor cores are not used (batched tiny factorisations).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef GPV_MAX_D
#define GPV_MAX_D 8
#endif

namespace gpv {

enum CovKind : int { COV_EXP = 0, COV_M15 = 1, COV_M25 = 2, COV_ESQE = 3, COV_GENERAL = 4 };

// Piecewise-polynomial table of the general-nu Matern (built per call, bessel_table.cuh).
// Variable: w = squared distance.  Interval index from the exponent and the top `sub_bits`
// mantissa bits of w; local variable v = mantissa(w) - centre in [-2^-(sub_bits+1), +..).
struct CovTable {
  const double* coef;   // [deg+1][nint], coefficient-major
  int nint;
  int idx0;             // (hi32(w) >> (20 - sub_bits)) - idx0 = interval index
  int sub_bits;
  int deg;
  double w_split;       // w > w_split : table holds exp(+s) * cov, multiply by exp(-s)
  // constants for the direct (out-of-table) evaluator, Temme / Steed CF2
  double nu, xmu, gam1, gam2, gampl, gammi, normcon;
  int nl;
};

struct UParams {
  int64_t nrows;          // rows of this shard
  int64_t row0;           // global index of the first row of the shard
  int p;                  // actual set size m+1 (<= P)
  int d;                  // spatial dimension
  const double* locs;     // [Nlocs][d] row-major
  const int32_t* nn;      // [nrows][p] row-major, 0-based ids, -1 = missing
  const uint64_t* cond;   // [nrows] bit j = (revCond[row, j] == TRUE)
  const double* nuggets;  // [Nlocs]
  double* out;            // U values or nullptr
  const int64_t* row_off; // packed offsets [nrows] or nullptr (-> row*p, zero filled)
  const double* zord;     // [n] or nullptr
  const int32_t* obsrank; // [Nlocs] rank among observed locs or -1
  int64_t skip_rows;      // global rows < skip_rows do not enter the likelihood sums
  double* partials;       // [gridDim.x][2] or nullptr
  unsigned long long* nfail;
  long long* first_fail;
  int cov;                // CovKind
  double c0;              // covariance at distance 0
  double c1, c2, c3, c4;  // kind-specific constants (host: make_cov_constants)
  double inv_range;       // general branch
  CovTable tab;
};

constexpr int kWarpsPerBlock = 8;

// --------------------------------------------------------------------------------------------
// covariance as a function of the SQUARED distance r2 (src/Matern.cpp, src/Esqe.cpp restated;
// every closed form returns exactly c0 at r2 == 0, as the reference's `dist == 0` branches do)
// --------------------------------------------------------------------------------------------
__device__ double matern_general_direct(double s, const CovTable& t);  // bessel_table.cuh
__device__ __forceinline__ double cov_general(double r2, const UParams& q);

template <int KIND>
__device__ __forceinline__ double cov_eval(double r2, const UParams& q) {
  if (KIND == COV_EXP) {          // Matern.cpp:38-39   sig2 exp(-d/range)          c1 = 1/range
    return q.c0 * exp(-sqrt(r2) * q.c1);
  } else if (KIND == COV_M15) {   // :51-52  sig2 (1+sqrt3 s) exp(-sqrt3 s)          c1 = sqrt3/range
    double t = sqrt(r2) * q.c1;
    return q.c0 * (1.0 + t) * exp(-t);
  } else if (KIND == COV_M25) {   // :66-68  sig2 exp(-t)(1 + t + t^2/3), t = sqrt5 s  c1 = sqrt5/range
    double t = sqrt(r2) * q.c1;
    return q.c0 * exp(-t) * fma(t, fma(t, 1.0 / 3.0, 1.0), 1.0);
  } else if (KIND == COV_ESQE) {  // Esqe.cpp:32-34  c4 = sig2_1, c1 = 1/r1, c2 = sig2_2, c3 = 1/r2^2
    return fma(q.c2, exp(-r2 * q.c3), q.c4 * exp(-sqrt(r2) * q.c1));
  } else {
    return cov_general(r2, q);
  }
}

template <int G, int P, int D>
struct GroupLayout {
  static constexpr int DD = (D > 0) ? D : GPV_MAX_D;
  static constexpr int LD = (P % 2 == 0) ? P + 1 : P;   // odd leading dimension
  static constexpr int kL = ((P * LD + 1) / 2) * 2;     // doubles, even => 16B aligned blocks
  static constexpr int kX = DD * G;                     // SoA coordinate staging
  static constexpr int kI = G / 2;                      // G int32 ids
  static constexpr int kDoubles = kL + kX + kI;
  static constexpr int kSetsPerWarp = 32 / G;
  static constexpr int kBytesPerBlock = kDoubles * 8 * kSetsPerWarp * kWarpsPerBlock;
};

// pair stage: lane gl evaluates the covariances (gl, gl+t mod P), t = 1..P/2
template <int KIND, int G, int P, int D>
__device__ __forceinline__ void pair_stage(const UParams& q, double* __restrict__ Ls,
                                           const double* __restrict__ xs, const double* xi,
                                           int gl, int npad, int d) {
  using LY = GroupLayout<G, P, D>;
  constexpr int LD = LY::LD;
#pragma unroll 1
  for (int t = 1; t <= P / 2; ++t) {
    int j = gl + t;
    if (j >= P) j -= P;
    double r2 = 0.0;
    if (D > 0) {
#pragma unroll
      for (int c = 0; c < LY::DD; ++c) {
        double dd = xi[c] - xs[c * G + j];
        r2 = fma(dd, dd, r2);
      }
    } else {
#pragma unroll
      for (int c = 0; c < GPV_MAX_D; ++c) {
        if (c < d) {
          double dd = xi[c] - xs[c * G + j];
          r2 = fma(dd, dd, r2);
        }
      }
    }
    double v = cov_eval<KIND>(r2, q);
    if (gl < npad || j < npad) v = 0.0;
    const bool active = (gl < P) && ((2 * t < P) || (gl < P / 2));
    const int hi = gl > j ? gl : j;
    const int lo = gl > j ? j : gl;
    if (active) Ls[hi * LD + lo] = v;
  }
}

template <int G, int P, int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
u_sets_kernel(const UParams q) {
  using LY = GroupLayout<G, P, D>;
  constexpr int LD = LY::LD;
  constexpr int SETS = LY::kSetsPerWarp;
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(P <= G && G <= 32, "set must fit its lane group");

  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / G;
  const int gl = lane % G;
  const int base = sub * G;
  const unsigned gmask = (G == 32) ? FULL : ((1u << G) - 1u);
  const int d = (D > 0) ? D : q.d;

  double* Ls = smem + (size_t)(warp * SETS + sub) * LY::kDoubles;
  double* xs = Ls + LY::kL;
  int* ids = reinterpret_cast<int*>(xs + LY::kX);

  double acc_quad = 0.0, acc_logd = 0.0;
  const int p = q.p;
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock * SETS;

  for (int64_t r0 = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * SETS; r0 < q.nrows; r0 += stride) {
    const int64_t row = r0 + sub;
    const bool row_ok = row < q.nrows;

    // ---- 1. ids, compaction (U_NZentries.cpp:41-45) ----------------------------------------
    int raw = -1;
    if (row_ok && gl < p) raw = q.nn[row * p + gl];
    const unsigned bal = __ballot_sync(FULL, raw >= 0);
    const unsigned bits = (bal >> base) & gmask;
    const int n0 = __popc(bits);
    const int npad = P - n0;
    // warp-uniform skip: nothing to factor anywhere in this warp
    if (bal == 0u) continue;
    {
      const int rank = __popc(bits & ((1u << gl) - 1u));
      if (gl < npad) ids[gl] = -1;
      __syncwarp();
      if (raw >= 0) ids[npad + rank] = raw;
      __syncwarp();
    }
    const int id = (gl < P) ? ids[gl] : -1;
    const bool real = id >= 0;
    const uint64_t cmask = row_ok ? q.cond[row] : 0ull;
    // compacted entry j reads revCond[row, p - n0 + j] (U_NZentries.cpp:47); local index gl = npad + j
    const int cpos = gl - (P - p);
    const bool condbit = real && ((cmask >> (cpos & 63)) & 1ull);

    // ---- 2. gather coordinates and nugget ----------------------------------------------------
    double xi[LY::DD];
    double diag = 1.0;
    if (real) {
      if (D == 2) {
        const double2 v = reinterpret_cast<const double2*>(q.locs)[id];
        xi[0] = v.x; xi[1] = v.y;
      } else {
#pragma unroll
        for (int c = 0; c < LY::DD; ++c) xi[c] = (c < d) ? q.locs[(int64_t)id * d + c] : 0.0;
      }
      // nug = nuggets[id] * (1 - revCond)   (U_NZentries.cpp:47; Inf * 0 = NaN kept on purpose)
      const double nug = q.nuggets[id] * (1.0 - (condbit ? 1.0 : 0.0));
      diag = q.c0 + nug;
    } else {
#pragma unroll
      for (int c = 0; c < LY::DD; ++c) xi[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < LY::DD; ++c) xs[c * G + gl] = xi[c];
    __syncwarp();

    // ---- 3. covariance pairs -> shared staging ------------------------------------------------
    switch (q.cov) {
      case COV_EXP: pair_stage<COV_EXP, G, P, D>(q, Ls, xs, xi, gl, npad, d); break;
      case COV_M15: pair_stage<COV_M15, G, P, D>(q, Ls, xs, xi, gl, npad, d); break;
      case COV_M25: pair_stage<COV_M25, G, P, D>(q, Ls, xs, xi, gl, npad, d); break;
      case COV_ESQE: pair_stage<COV_ESQE, G, P, D>(q, Ls, xs, xi, gl, npad, d); break;
      default: pair_stage<COV_GENERAL, G, P, D>(q, Ls, xs, xi, gl, npad, d); break;
    }
    if (gl < P) Ls[gl * LD + gl] = diag;
    __syncwarp();

    // ---- 4. row of the lower triangle into registers -------------------------------------------
    double a[P];
    {
      const int rr = gl < P ? gl : P - 1;
#pragma unroll
      for (int j = 0; j < P; ++j) a[j] = Ls[rr * LD + j];   // j > rr: stale garbage, never used
    }
    __syncwarp();

    // ---- 5. right-looking Cholesky (chol(covmat,"upper"), U_NZentries.cpp:61) --------------------
    bool fail = false;
    double myinv = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double akk = __shfl_sync(FULL, a[k], base + k);
      fail = fail || !(akk > 0.0);          // dpotrf: leading minor not positive definite (or NaN)
      const double inv = rsqrt(akk);        // rsqrt(+Inf) = 0 : an Inf nugget decouples that neighbour
      const double l = a[k] * inv;          // lane r >= k: L[r][k]; lane k: sqrt(akk)
      if (gl == k) myinv = inv;
      if (gl < P) Ls[k * LD + gl] = l;      // column k of L, contiguous
      __syncwarp();
      // trailing update of my row: a[j] -= L[r][k] * L[j][k], j = k+1..P-1 (garbage for j > r)
      constexpr int dummy = 0; (void)dummy;
      int j = k + 1;
      if (j < P) {                          // (k*LD + j) is odd here: one scalar broadcast load
        a[j] = fma(-l, Ls[k * LD + j], a[j]);
        ++j;
      }
#pragma unroll
      for (; j + 1 < P; j += 2) {           // (k*LD + j) even: 16B broadcast loads
        const double2 lj = *reinterpret_cast<const double2*>(&Ls[k * LD + j]);
        a[j] = fma(-l, lj.x, a[j]);
        a[j + 1] = fma(-l, lj.y, a[j + 1]);
      }
      if (j < P) a[j] = fma(-l, Ls[k * LD + j], a[j]);
    }

    // ---- 6. x = L^{-T} e_P  (solve(R, onevec), U_NZentries.cpp:62) --------------------------------
    double s = 0.0;
    double x = 0.0;
#pragma unroll
    for (int j = P - 1; j >= 1; --j) {
      const double cand = (j == P - 1) ? myinv : -s * myinv;   // x_j on lane j
      const double xj = __shfl_sync(FULL, cand, base + j);
      if (gl == j) x = xj;
      const int rr = gl < j ? gl : 0;
      const double lji = Ls[rr * LD + j];                      // L[j][gl]
      if (gl < j) s = fma(lji, xj, s);
    }
    if (gl == 0) x = (P == 1) ? myinv : -s * myinv;
    if (fail) x = 0.0;                                         // row stays zero (:64-66)

    // ---- 7. outputs ---------------------------------------------------------------------------------
    if (fail && row_ok && gl == 0 && n0 > 0) {
      atomicAdd(q.nfail, 1ull);
      atomicMin(q.first_fail, (long long)(q.row0 + row));
    }
    if (q.out != nullptr && row_ok) {
      if (q.row_off != nullptr) {
        if (real) q.out[q.row_off[row] + (gl - npad)] = x;
      } else {
        double* o = q.out + row * (int64_t)p;
        if (real) o[gl - npad] = x;
        else if (gl < P && n0 + gl < p) o[n0 + gl] = 0.0;     // zero fill beyond n0 (:33)
      }
    }
    if (q.partials != nullptr) {
      // quadform: (sum_{j: revCond = 0} x_j * zord[obsrank(id_j)])^2 ; logdet: log x_self
      double t = 0.0;
      if (real && !condbit) {
        const int orank = q.obsrank[id];
        if (orank >= 0) t = x * q.zord[orank];
      }
#pragma unroll
      for (int o = G / 2; o >= 1; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
      const double xs_self = __shfl_sync(FULL, x, base + P - 1);
      if (gl == 0 && row_ok && n0 > 0 && (q.row0 + row) >= q.skip_rows) {
        acc_quad += t * t;
        acc_logd += log(xs_self);
      }
    }
    __syncwarp();
  }

  // ---- deterministic block reduction of the likelihood partial sums ---------------------------------
  if (q.partials != nullptr) {
    __shared__ double red[kWarpsPerBlock][2];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      acc_quad += __shfl_xor_sync(FULL, acc_quad, o);
      acc_logd += __shfl_xor_sync(FULL, acc_logd, o);
    }
    if (lane == 0) { red[warp][0] = acc_quad; red[warp][1] = acc_logd; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a0 = 0.0, a1 = 0.0;
      for (int w = 0; w < kWarpsPerBlock; ++w) { a0 += red[w][0]; a1 += red[w][1]; }
      q.partials[2 * blockIdx.x] = a0;
      q.partials[2 * blockIdx.x + 1] = a1;
    }
  }
}

}  // namespace gpv
