// u_quad.cuh -- the set kernel for 24 < P <= 32 (m = 24..31), "quad-folded": a lane group of EIGHT
// lanes owns one conditioning set, so a warp instruction serves FOUR sets, and every lane keeps FOUR
// rows of the lower triangle in registers.
//
// Same path as u_kernels.cuh (reference src/U_NZentries.cpp:39-69, R/vecchia_likelihood.R:74-91) and
// the same arithmetic in the same order; only the mapping of rows to lanes differs.  Why it exists
// (profiles/r01_u_sets_closed_*): the two-rows-per-lane kernel is bound by the L1/shared data pipe
// (85 % of its wavefront rate), and a shared-memory load costs one wavefront per distinct 8-byte word
// per warp instruction whether it serves 2 sets or 4 (tools/microbench/lds3.cu).  Broadcasting a
// column of L to four sets per wavefront halves that traffic per set, four rows per lane double the
// FMAs fed by each broadcast word, and a warp's instruction stream carries four independent
// covariance chains per iteration instead of two.
//
// Row ownership: lane q (0..7) holds rows q, 15-q, 16+q and 31-q ("bands" 0..3; rows >= P do not
// exist).  The lengths of a lane's rows add up to 66 for every q, so the triangular load is balanced
// like in the two-row fold.  Register arrays are sized by the longest row of a band: 8 + 16 + 24 + P.
#pragma once
#include "u_kernels.cuh"

namespace gpv {

template <int P, int D>
struct QuadLayout {
  static constexpr int G = 8;
  static constexpr int DD = (D > 0) ? D : GPV_MAX_D;
  static constexpr int S0 = 8, S1 = 16, S2 = 24, S3 = P;   // static row lengths of the bands
  static constexpr int kScratch = tri_col(P, P);           // per-lane dump slots for inactive pair stores
  static constexpr int kBuf = ((tri_col(P, P) + G + 1) / 2) * 2;
  static constexpr int kT = P / 2;                         // pair-stage iterations
  static constexpr int PX = 32;                            // coordinate row stride
  static constexpr int kX = DD * PX;
  static constexpr int kNug = PX;
  static constexpr int kZ = PX;
  static constexpr int kI = 16;                            // 32 int32 compacted ids
  static constexpr int kRawI = 16;                         // 32 int32 raw ids of a row (as stored)
  static constexpr int kMeta = 2;                          // cond mask (8 B), row (4 B), pad
  static constexpr int kStage = kX + kNug + kZ + kI + kRawI + kMeta;
  static constexpr int kOffNug = kX, kOffZ = kX + kNug, kOffIds = kX + kNug + kZ,
                       kOffRaw = kX + kNug + kZ + kI, kOffMeta = kX + kNug + kZ + kI + kRawI;
  static constexpr int kSetsPerWarp = 4;
  static constexpr int kRaw = kBuf + 2 * kStage;
  // every set starts on a 128-byte line plus a skew of {0, 64, 32, 96} bytes: the 64-byte row
  // segments of the two sets of a half-warp tile one line, and the four broadcast words of a warp
  // instruction fall into different banks
  static constexpr int kDoubles = ((kRaw + 15) / 16) * 16 + 16;
  static constexpr int kBytesPerBlock = kDoubles * 8 * kSetsPerWarp * kWarpsPerBlock;
  static_assert(P > 24 && P <= 32, "quad-folded kernel: 24 < P <= 32");
};

#ifndef GPV_QUAD_SOLVE_BLOCK
#define GPV_QUAD_SOLVE_BLOCK 1
#endif
constexpr int kQuadSolveBlock = GPV_QUAD_SOLVE_BLOCK;
#ifndef GPV_QUAD_FACT
#define GPV_QUAD_FACT 1
#endif
constexpr bool kQuadUnscaled = (GPV_QUAD_FACT == 1);   // columns of L kept as a[r][k] = L[r][k] d_k in shared memory   // rows per shuffle round of the back substitution

__host__ __device__ constexpr int quad_band(int r) { return r >> 3; }
__host__ __device__ constexpr int quad_owner(int r) {   // lane (within the group) that holds row r
  return ((r >> 3) & 1) ? (8 * (r >> 3) + 7 - r) : (r - 8 * (r >> 3));
}

// Per (iteration t, lane q): where the four pairs (i, i + t mod P) of the lane's points go in the packed
// staged triangle (byte offsets, 16 bits each; inactive combinations point at the lane's dump slot) and
// which staged point is the partner (index i + t mod P, 16 bits each).
template <int P, int D>
__device__ __forceinline__ void build_store_table_quad(uint4* __restrict__ stab) {
  using LY = QuadLayout<P, D>;
  for (int idx = threadIdx.x; idx < LY::kT * 8; idx += blockDim.x) {
    const int t = idx / 8 + 1, q = idx % 8;
    const bool full = (2 * t < P);                       // even P: t = P/2 is covered by i < P/2 only
    unsigned off[4], par[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = (b == 0) ? q : (b == 1) ? 15 - q : (b == 2) ? 16 + q : 31 - q;
      const int ic = i < P ? i : P - 1;
      int j = ic + t; if (j >= P) j -= P;
      par[b] = j;
      off[b] = LY::kScratch + q;                         // own dump slot: no write-write race
      if (i < P && (full || i < P / 2)) {
        const int a = i > j ? i : j, bb = i > j ? j : i;
        off[b] = tri_col(bb, P) + a - bb;
      }
      off[b] *= 8;
    }
    stab[idx] = make_uint4(off[0] | (off[1] << 16), off[2] | (off[3] << 16), par[0] | (par[1] << 16), par[2] | (par[3] << 16));
  }
}

// pair stage: point i evaluates the covariances (i, i + t mod P), t = 1..P/2; a lane carries its
// four points through the same iteration (four independent chains).
template <int KIND, int P, int D, class C>
__device__ __forceinline__ void pair_stage_quad_impl(const C& q, double* __restrict__ As,
                                                     const double* __restrict__ xs,
                                                     const double (&x)[4][QuadLayout<P, D>::DD], const int (&rc)[4],
                                                     int gl, const uint4* __restrict__ stab,
                                                     const double* __restrict__ etab, int d) {
  using LY = QuadLayout<P, D>;
  const uint4* stab_lane = stab + gl;
  char* Asb = reinterpret_cast<char*>(As);
  const double guard = (KIND == COV_GENERAL) ? 0.0 : kMathC[7];
#pragma unroll 1
  for (int t = 1; t <= LY::kT; ++t) {
    const uint4 offs = stab_lane[(t - 1) * 8];
    const int jp[4] = {(int)(offs.z & 0xffffu), (int)(offs.z >> 16), (int)(offs.w & 0xffffu), (int)(offs.w >> 16)};
    double r2[4], v[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[b] = pair_r2<D>(xs, LY::PX, x[b], jp[b], d, guard);
    if constexpr (KIND == COV_GENERAL) {
      int idx[4];
      bool sp = false;
#pragma unroll
      for (int b = 0; b < 4; ++b) sp |= cov_general_special(r2[b], q.tab, &idx[b]);
      if (__any_sync(0xffffffffu, sp)) {
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = cov_general_slow(r2[b], q, etab);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = cov_general_fast(r2[b], idx[b], q.tab);
      }
    } else {
      cov_eval_n<KIND, 4>(r2, v, q, etab);
    }
    *reinterpret_cast<double*>(Asb + (offs.x & 0xffffu)) = v[0];
    *reinterpret_cast<double*>(Asb + (offs.x >> 16)) = v[1];
    *reinterpret_cast<double*>(Asb + (offs.y & 0xffffu)) = v[2];
    *reinterpret_cast<double*>(Asb + (offs.y >> 16)) = v[3];
  }
}
// Closed forms, D = 2: compiled as a function of its own.  Inlined into the 255-register kernel,
// ptxas schedules the four chains one after the other on shared registers; as a separate function
// it interleaves them (profiles/).  Everything is passed by value so that neither the kernel
// parameters nor the caller's arrays are forced into local memory.
template <int KIND, int P>
static __device__ __noinline__ void pair_stage_quad_d2(CovConsts cc, double* __restrict__ As,
                                                       const double* __restrict__ xs, const double (&x)[4][2],
                                                       const int (&rc)[4], int gl, const uint4* __restrict__ stab,
                                                       const double* __restrict__ etab) {
  pair_stage_quad_impl<KIND, P, 2, CovConsts>(cc, As, xs, x, rc, gl, stab, etab, 2);
}
template <int KIND, int P, int D>
__device__ __forceinline__ void pair_stage_quad(const UParams& q, double* __restrict__ As,
                                                const double* __restrict__ xs,
                                                const double (&x)[4][QuadLayout<P, D>::DD], const int (&rc)[4],
                                                int gl, const uint4* __restrict__ stab,
                                                const double* __restrict__ etab, int d) {
  if constexpr (KIND != COV_GENERAL && D == 2) {
    const CovConsts cc = {q.c0, q.c1, q.c2, q.c3, q.c4};
    pair_stage_quad_d2<KIND, P>(cc, As, xs, x, rc, gl, stab, etab);
  } else {
    pair_stage_quad_impl<KIND, P, D, UParams>(q, As, xs, x, rc, gl, stab, etab, d);
  }
}

template <int P, int D, bool GENERAL>
__global__ void __launch_bounds__(kThreadsPerBlock, 2)
u_quad_kernel(const UParams q) {
  using LY = QuadLayout<P, D>;
  constexpr int G = 8, SETS = 4, PX = LY::PX;
  constexpr int S0 = LY::S0, S1 = LY::S1, S2 = LY::S2, S3 = LY::S3;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int kSelfLane = 32 - P;                  // lane that holds row P-1 (band 3)

  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane >> 3;
  const int gl = lane & 7;
  const int base = sub * G;
  const int d = (D > 0) ? D : q.d;
  const int p = q.p;

  double* buf = smem + (size_t)(warp * SETS + sub) * LY::kDoubles + ((sub & 1) * 8 + (sub >> 1) * 4);
  double* stage0 = buf + LY::kBuf;

  __shared__ uint4 stab[LY::kT * 8];
  __shared__ double etab[64];
  build_store_table_quad<P, D>(stab);
  if (threadIdx.x < 64) etab[threadIdx.x] = kExp2Tab[threadIdx.x];
  __syncthreads();

  // my four rows; a row that does not exist (31 - gl >= P) is clamped to P-1 and masked
  const int r3 = 31 - gl;
  const bool v3 = r3 < P;
  const int rc[4] = {gl, 15 - gl, 16 + gl, v3 ? r3 : P - 1};
  const bool vb[4] = {true, true, true, v3};

  double acc_quad = 0.0, acc_logd = 0.0, acc_qden = 0.0, acc_lden = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock * SETS;
  const int64_t first = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * SETS;

  // ---- input pipeline (cp.async; same two-stage scheme as u_sets_kernel) ---------------------------
  auto fetch_raw = [&](int64_t sidx, double* st) {
    int* raw = reinterpret_cast<int*>(st + LY::kOffRaw);
    double* meta = st + LY::kOffMeta;
    int* metai = reinterpret_cast<int*>(meta + 1);
    if (sidx < q.nsets) {
      const int32_t* nnr = q.nn + sidx * (int64_t)p;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int e = gl + 8 * c;
        if (e < p) __pipeline_memcpy_async(raw + e, nnr + e, 4); else raw[e] = -1;
      }
      if (gl == 0) {
        __pipeline_memcpy_async(meta, q.cond + sidx, 8);
        if (q.rowmap != nullptr) __pipeline_memcpy_async(metai, q.rowmap + sidx, 4);
        else metai[0] = (int)(q.set_base + sidx);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) raw[gl + 8 * c] = -1;
      if (gl == 0) { reinterpret_cast<unsigned long long*>(meta)[0] = 0ull; metai[0] = -1; }
    }
  };
  // compaction of the raw ids (U_NZentries.cpp:41-45; entries gl + 8c of the row), then coordinates,
  // nuggets and z of my four points -> stage st.  Returns n0.
  auto gather = [&](double* st) -> int {
    const int* raw = reinterpret_cast<const int*>(st + LY::kOffRaw);
    int* ids = reinterpret_cast<int*>(st + LY::kOffIds);
    double* xs = st;
    double* nug = st + LY::kOffNug;
    int rawv[4];
    unsigned bal[4];
    int n0 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      rawv[c] = raw[gl + 8 * c];
      bal[c] = (__ballot_sync(FULL, rawv[c] >= 0) >> base) & 0xffu;
      n0 += __popc(bal[c]);
    }
    const int npad = P - n0;
    const unsigned below = (1u << gl) - 1u;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (gl + 8 * c < npad) ids[gl + 8 * c] = -1;
    __syncwarp();
    int pre = npad;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (rawv[c] >= 0) ids[pre + __popc(bal[c] & below)] = rawv[c];
      pre += __popc(bal[c]);
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (!vb[b]) continue;
      const int r = rc[b];
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
    const uint64_t cmask = reinterpret_cast<const unsigned long long*>(st + LY::kOffMeta)[0];
    const int row = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[0];
    const bool row_ok = row >= 0;
    const int npad = P - n0;
    const int n0_next = gather(stn);                       // set i+1: ids now, coordinates in flight
    __syncwarp();
    fetch_raw(s0 + 2 * stride + sub, st);                  // set i+2: ids in flight
    __pipeline_commit();

    // ---- 1./2. my four points of the current set ----------------------------------------------------
    const double* xs = st;
    const int* ids = reinterpret_cast<const int*>(st + LY::kOffIds);
    const double* nugs = st + LY::kOffNug;
    double x[4][LY::DD];
    double dg[4];
    int id[4];
    bool cd[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      id[b] = vb[b] ? ids[rc[b]] : -1;
      if (D == 2) {
        const double2 vv = reinterpret_cast<const double2*>(xs)[rc[b]];
        x[b][0] = vv.x; x[b][1] = vv.y;
      } else {
#pragma unroll
        for (int c = 0; c < LY::DD; ++c) x[b][c] = (c < d) ? xs[c * PX + rc[b]] : 0.0;
      }
      dg[b] = 1.0;
      cd[b] = false;
      if (id[b] >= 0) {
        // compacted entry j reads revCond[row, p - n0 + j] (:47); local index = npad + j
        cd[b] = (cmask >> ((rc[b] - (P - p)) & 63)) & 1ull;
        dg[b] = q.c0 + clamp_nugget(nugs[rc[b]] * (1.0 - (cd[b] ? 1.0 : 0.0)));   // Inf * 0 = NaN kept
      }
    }

    // ---- 3. covariance pairs -> shared staging (packed lower triangle) -------------------------------
    if (GENERAL) {
      pair_stage_quad<COV_GENERAL, P, D>(q, buf, xs, x, rc, gl, stab, etab, d);
    } else {
      switch (q.cov) {
        case COV_EXP: pair_stage_quad<COV_EXP, P, D>(q, buf, xs, x, rc, gl, stab, etab, d); break;
        case COV_M15: pair_stage_quad<COV_M15, P, D>(q, buf, xs, x, rc, gl, stab, etab, d); break;
        case COV_M25: pair_stage_quad<COV_M25, P, D>(q, buf, xs, x, rc, gl, stab, etab, d); break;
        default: pair_stage_quad<COV_ESQE, P, D>(q, buf, xs, x, rc, gl, stab, etab, d); break;
      }
    }
    if (__any_sync(FULL, npad > 0)) {
      // padding occupies the leading indices: zero columns 0..npad-1 of the staged triangle
      __syncwarp();
      for (int j = 0; j < npad; ++j) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (vb[b] && rc[b] > j) buf[tri_col(j, P) + rc[b] - j] = 0.0;
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (vb[b]) buf[tri_col(rc[b], P)] = dg[b];
    __syncwarp();

    // ---- 4. my four rows of the lower triangle into registers (entries beyond the diagonal: stale,
    // never used) ---------------------------------------------------------------------------------------
    double a0[S0], a1[S1], a2[S2], a3[S3];
#pragma unroll
    for (int j = 0; j < S0; ++j) a0[j] = buf[tri_col(j, P) + rc[0] - j];
#pragma unroll
    for (int j = 0; j < S1; ++j) a1[j] = buf[tri_col(j, P) + rc[1] - j];
#pragma unroll
    for (int j = 0; j < S2; ++j) a2[j] = buf[tri_col(j, P) + rc[2] - j];
#pragma unroll
    for (int j = 0; j < S3; ++j) a3[j] = buf[tri_col(j, P) + rc[3] - j];
    __syncwarp();

#if GPV_QUAD_FACT == 1
    // ---- 5. right-looking LDL^T (square-root-free Cholesky; chol(covmat,"upper"), U_NZentries.cpp:61)
    // Sigma = L D L^T.  Columns are published to shared memory UNSCALED (a[r][k] = L[r][k] d_k), as soon
    // as they are final: column k+1 is the first thing step k updates, and it is stored before the rest
    // of the trailing update is issued.  Every lane then reads the pivot a[k][k] and the column back as
    // group-broadcast loads and scales its own four multipliers m = a[r][k] / d_k.  No shuffle and no
    // reciprocal sit between a column becoming final and its publication, which is what bounds a step
    // when only two warps share a scheduler.  Column 0 is the staged matrix itself.  1 / d_k replaces
    // the pivot in its diagonal slot one step later (nobody reads that slot between), for the sweep.
    bool fail = false;
    double dlast = 1.0;
    double inv_prev = 0.0;
#define GPV_Q_UPD(J, W)                                                                   \
    do {                                                                                  \
      if ((J) < S0) a0[(J) < S0 ? (J) : 0] = fma(-m0, (W), a0[(J) < S0 ? (J) : 0]);         \
      if ((J) < S1) a1[(J) < S1 ? (J) : 0] = fma(-m1, (W), a1[(J) < S1 ? (J) : 0]);         \
      if ((J) < S2) a2[(J) < S2 ? (J) : 0] = fma(-m2, (W), a2[(J) < S2 ? (J) : 0]);         \
      a3[(J)] = fma(-m3, (W), a3[(J)]);                                                   \
    } while (0)
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const int ck = tri_col(k, P) - k;     // a[r][k] at buf[ck + r], r >= k
      double akk, wa = 0.0, wb = 0.0;       // pivot, a[k+1][k], a[k+2][k]
      int j;                                // first column of the paired loop
      if (k == P - 1) {
        akk = buf[ck + k];
        j = P;
      } else if (((ck + k) & 1) == 0) {
        const double2 l2 = *reinterpret_cast<const double2*>(&buf[ck + k]);
        akk = l2.x; wa = l2.y; j = k + 2;
      } else {
        akk = buf[ck + k];
        if (k + 2 < P) {
          const double2 l2 = *reinterpret_cast<const double2*>(&buf[ck + k + 1]);
          wa = l2.x; wb = l2.y; j = k + 3;
        } else {
          wa = buf[ck + k + 1]; j = k + 2;
        }
      }
      // positive, normal, finite -- dpotrf's `ajj <= 0 || isnan(ajj)` test on the integer pipe
      fail = fail || ((unsigned)(__double2hiint(akk) - 0x00100000) >= 0x7fe00000u);
      if (k >= 1 && gl == quad_owner(k - 1)) buf[tri_col(k - 1, P)] = inv_prev;
      if (k == P - 1) { dlast = akk; break; }
      const double inv = rcp_pos(akk);      // an Inf nugget arrives here as 1e300 (clamp_nugget)
      inv_prev = inv;
      const double m0 = (k < S0) ? a0[k < S0 ? k : 0] * inv : 0.0;
      const double m1 = (k < S1) ? a1[k < S1 ? k : 0] * inv : 0.0;
      const double m2 = (k < S2) ? a2[k < S2 ? k : 0] * inv : 0.0;
      const double m3 = a3[k] * inv;
      // column k+1: update, publish
      GPV_Q_UPD(k + 1, wa);
      {
        const int cn = tri_col(k + 1, P) - (k + 1);
        if (k + 1 < S0 && rc[0] >= k + 1) buf[cn + rc[0]] = a0[k + 1 < S0 ? k + 1 : 0];
        if (k + 1 < S1 && rc[1] >= k + 1) buf[cn + rc[1]] = a1[k + 1 < S1 ? k + 1 : 0];
        if (k + 1 < S2 && rc[2] >= k + 1) buf[cn + rc[2]] = a2[k + 1 < S2 ? k + 1 : 0];
        if (v3 && rc[3] >= k + 1) buf[cn + rc[3]] = a3[k + 1];
      }
      // the rest of the trailing update: a[r][j] -= m_r a[j][k]
      if (j == k + 3) GPV_Q_UPD(k + 2, wb);
#pragma unroll
      for (; j + 1 < P; j += 2) {           // 16-byte aligned broadcast loads for (j, j+1)
        const double2 l2 = *reinterpret_cast<const double2*>(&buf[ck + j]);
        GPV_Q_UPD(j, l2.x);
        GPV_Q_UPD(j + 1, l2.y);
      }
      if (j < P) {
        const double l1 = buf[ck + j];
        GPV_Q_UPD(j, l1);
      }
      __syncwarp();
    }
#undef GPV_Q_UPD
    __syncwarp();

#else
    // ---- 5. right-looking LDL^T (see u_sets_kernel step 5): columns published divided by the pivot ------
    bool fail = false;
    double dlast = 1.0;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const int own = base + quad_owner(k);
      const double akk = (k < S0)   ? __shfl_sync(FULL, a0[k < S0 ? k : 0], own)
                         : (k < S1) ? __shfl_sync(FULL, a1[k < S1 ? k : 0], own)
                         : (k < S2) ? __shfl_sync(FULL, a2[k < S2 ? k : 0], own)
                                    : __shfl_sync(FULL, a3[k], own);
      fail = fail || ((unsigned)(__double2hiint(akk) - 0x00100000) >= 0x7fe00000u);
      if (k == P - 1) { dlast = akk; break; }
      const double inv = rcp_pos(akk);
      const int ck = tri_col(k, P) - k;     // L[r][k] at buf[ck + r]
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      if (k < S0) { c0 = a0[k < S0 ? k : 0]; if (rc[0] >= k) buf[ck + rc[0]] = c0 * inv; }
      if (k < S1) { c1 = a1[k < S1 ? k : 0]; if (rc[1] >= k) buf[ck + rc[1]] = c1 * inv; }
      if (k < S2) { c2 = a2[k < S2 ? k : 0]; if (rc[2] >= k) buf[ck + rc[2]] = c2 * inv; }
      const double c3 = a3[k];
      if (v3 && rc[3] >= k) buf[ck + rc[3]] = c3 * inv;
      __syncwarp();
      int j = k + 1;
      if (j < P && ((ck + j) & 1) != 0) {   // odd offset: one scalar broadcast load first
        const double l1 = buf[ck + j];
        if (j < S0) a0[j < S0 ? j : 0] = fma(-c0, l1, a0[j < S0 ? j : 0]);
        if (j < S1) a1[j < S1 ? j : 0] = fma(-c1, l1, a1[j < S1 ? j : 0]);
        if (j < S2) a2[j < S2 ? j : 0] = fma(-c2, l1, a2[j < S2 ? j : 0]);
        a3[j] = fma(-c3, l1, a3[j]);
        ++j;
      }
#pragma unroll
      for (; j + 1 < P; j += 2) {           // 16-byte aligned broadcast loads for (j, j+1)
        const double2 l2 = *reinterpret_cast<const double2*>(&buf[ck + j]);
        if (j < S0) a0[j < S0 ? j : 0] = fma(-c0, l2.x, a0[j < S0 ? j : 0]);
        if (j + 1 < S0) a0[j + 1 < S0 ? j + 1 : 0] = fma(-c0, l2.y, a0[j + 1 < S0 ? j + 1 : 0]);
        if (j < S1) a1[j < S1 ? j : 0] = fma(-c1, l2.x, a1[j < S1 ? j : 0]);
        if (j + 1 < S1) a1[j + 1 < S1 ? j + 1 : 0] = fma(-c1, l2.y, a1[j + 1 < S1 ? j + 1 : 0]);
        if (j < S2) a2[j < S2 ? j : 0] = fma(-c2, l2.x, a2[j < S2 ? j : 0]);
        if (j + 1 < S2) a2[j + 1 < S2 ? j + 1 : 0] = fma(-c2, l2.y, a2[j + 1 < S2 ? j + 1 : 0]);
        a3[j] = fma(-c3, l2.x, a3[j]);
        a3[j + 1] = fma(-c3, l2.y, a3[j + 1]);
      }
      if (j < P) {
        const double l1 = buf[ck + j];
        if (j < S0) a0[j < S0 ? j : 0] = fma(-c0, l1, a0[j < S0 ? j : 0]);
        if (j < S1) a1[j < S1 ? j : 0] = fma(-c1, l1, a1[j < S1 ? j : 0]);
        if (j < S2) a2[j < S2 ? j : 0] = fma(-c2, l1, a2[j < S2 ? j : 0]);
        a3[j] = fma(-c3, l1, a3[j]);
      }
    }
    __syncwarp();

#endif
    // ---- 6. x = L^{-T} e_P / sqrt(d_P)  (solve(R, onevec), U_NZentries.cpp:62): unit-triangular column
    // sweep on t = -y, t_r = -sum_{j > r} L[j][r] t_j with L[j][r] = a[j][r] / d_r rebuilt from the
    // unscaled column and the reciprocal in its diagonal slot; the unit right-hand side enters as
    // t_{P-1} = -1.  kQuadSolveBlock rows per round: their partial sums are shuffled together, every lane
    // finishes the block's small triangle redundantly from broadcast loads, then applies the block to its
    // own rows; all shared-memory operands of a round are fetched before its shuffles.
    double s[4] = {0.0, 0.0, 0.0, (gl == kSelfLane) ? -1.0 : 0.0};
    int cb[4];
    double invd[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { cb[b] = tri_col(rc[b], P) - rc[b]; invd[b] = kQuadUnscaled ? buf[tri_col(rc[b], P)] : 1.0; }
    constexpr int B = kQuadSolveBlock;
    if constexpr (B == 1) {
#pragma unroll
      for (int j = P - 1; j >= 1; --j) {
        const double tj = __shfl_sync(FULL, s[quad_band(j)], base + quad_owner(j));
        // L[j][r] = a[j][r] / d_r, a[j][r] at buf[tri_col(r) - r + j]; band b holds rows 8b..8b+7
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (j > 8 * b && rc[b] < j) s[b] = fma(-(kQuadUnscaled ? buf[cb[b] + j] * invd[b] : buf[cb[b] + j]), tj, s[b]);
      }
    } else {
#pragma unroll
    for (int j = P - 1; j >= 1; j -= B) {
      double mm[B][4], lt[B][B], t[B];
#pragma unroll
      for (int bb = 0; bb < B; ++bb) {
        const int r = j - bb;
#pragma unroll
        for (int b = 0; b < 4; ++b) mm[bb][b] = (r >= 1 && r > 8 * b) ? buf[cb[b] + r] * invd[b] : 0.0;
      }
#pragma unroll
      for (int bb = 1; bb < B; ++bb) {
        const int rb = j - bb;
#pragma unroll
        for (int a = 0; a < bb; ++a)            // L[j-a][rb]
          lt[a][bb] = (rb >= 1) ? buf[tri_col(rb > 0 ? rb : 0, P) - rb + (j - a)] * (kQuadUnscaled ? buf[tri_col(rb > 0 ? rb : 0, P)] : 1.0) : 0.0;
      }
#pragma unroll
      for (int bb = 0; bb < B; ++bb) {
        const int r = (j - bb >= 1) ? j - bb : 1;
        t[bb] = __shfl_sync(FULL, s[quad_band(r)], base + quad_owner(r));
      }
#pragma unroll
      for (int bb = 1; bb < B; ++bb) {
#pragma unroll
        for (int a = 0; a < bb; ++a)
          if (j - bb >= 1) t[bb] = fma(-lt[a][bb], t[a], t[bb]);
      }
#pragma unroll
      for (int bb = 0; bb < B; ++bb) {
        const int r = j - bb;
        if (r < 1) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (r > 8 * b && rc[b] < r) s[b] = fma(-mm[bb][b], t[bb], s[b]);
      }
    }
    }
    const double rs = rsqrt_pos(dlast);
    double xo[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) xo[b] = fail ? 0.0 : -s[b] * rs;   // failed row stays zero (:64-66)

    // ---- 7. outputs ------------------------------------------------------------------------------------
    if (fail && row_ok && gl == 0 && n0 > 0) {
      atomicAdd(q.nfail, 1ull);
      atomicMin(q.first_fail, (long long)(q.row0 + row));
    }
    if (q.out != nullptr && row_ok) {
      if (q.row_off != nullptr) {
        double* o = q.out + q.row_off[row];
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (id[b] >= 0) o[rc[b] - npad] = xo[b];
      } else {
        double* o = q.out + (int64_t)row * p;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (id[b] >= 0) o[rc[b] - npad] = xo[b];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gl + 8 * c >= n0 && gl + 8 * c < p) o[gl + 8 * c] = 0.0;   // zero fill beyond n0 (:33)
      }
    }
    if (q.partials != nullptr) {
      const double* zst = st + LY::kOffZ;
      double t = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (id[b] >= 0 && !cd[b]) t = fma(xo[b], zst[rc[b]], t);
      t += __shfl_xor_sync(FULL, t, 4);
      t += __shfl_xor_sync(FULL, t, 2);
      t += __shfl_xor_sync(FULL, t, 1);
      const double xself = __shfl_sync(FULL, xo[3], base + kSelfLane);
      if (gl == 0 && row_ok && n0 > 0 && (q.row0 + row) >= q.skip_rows) {
        acc_quad += t * t;
        acc_logd += log(xself);
        if (q.full_z) {
          const double tau = nugs[P - 1], zk = zst[P - 1];
          const double w = fma(xself, xself, 1.0 / tau);
          const double z2 = fma(xself, t, -zk / tau);
          acc_qden += z2 * z2 / w;
          acc_lden += log(w);
        }
      }
    }
    __pipeline_wait_prior(0);
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
      double a = 0.0;
      for (int w = 0; w < kWarpsPerBlock; ++w) a += red[w][threadIdx.x];
      q.partials[4 * blockIdx.x + threadIdx.x] = a;
    }
  }
}

}  // namespace gpv
