// u_band.cuh -- the set kernel for 2G < P <= 4G, "band-folded": a lane group of G lanes owns one
// conditioning set and every lane keeps THREE or FOUR rows of the lower triangle in registers.
//   G = 8  : 16 < P <= 32 (m = 16..31), four sets per warp instruction
//   G = 16 : 32 < P <= 48 (m = 32..47), two sets per warp instruction (three bands)
//
// Same path as u_kernels.cuh (reference src/U_NZentries.cpp:39-69, R/vecchia_likelihood.R:74-91);
// what differs is the mapping of rows to lanes and the order of work inside the factorisation.  Why it
// exists (profiles/r01_u_sets_closed_*): the two-rows-per-lane kernel is bound by the L1/shared data
// pipe (85 % of its wavefront rate).  A shared-memory load costs one wavefront per distinct 8-byte
// word per warp instruction whether it serves one set or four (tools/microbench/lds3.cu), so
// broadcasting a column of L to twice as many sets per wavefront halves that traffic per set, twice
// as many rows per lane double the FMAs fed by each broadcast word, and a warp's instruction stream
// carries three or four independent covariance chains per iteration instead of two.
//
// Row ownership: the rows are cut into bands of G; lane q holds row bG + q of an even band b and row
// (b+1)G - 1 - q of an odd one (rows >= P do not exist).  With four bands the lengths of a lane's
// rows add up to the same number for every q, so the triangular load is balanced like in the
// two-row fold.  Register arrays are sized by the longest row of a band: G, 2G, 3G (or P), P.
#pragma once
#include "u_kernels.cuh"

namespace gpv {

#ifndef GPV_BAND_STAGGER
#define GPV_BAND_STAGGER 6000   // cycles; measured: esqe 2.68 -> 2.58 ms, the Matern forms unchanged (profiles/r02_variants.log)
#endif
#ifndef GPV_BAND_MINB3
#define GPV_BAND_MINB3 3
#endif
template <int G, int P, int D>
struct BandLayout {
  static constexpr int NB = (P + G - 1) / G;               // bands: 3 or 4
  static constexpr int DD = (D > 0) ? D : GPV_MAX_D;
  static constexpr int S0 = G, S1 = 2 * G, S2 = (3 * G < P) ? 3 * G : P, S3 = (NB > 3) ? P : 1;
  static constexpr int kScratch = tri_col(P, P);           // per-lane dump slots for inactive pair stores
  static constexpr int kBuf = ((tri_col(P, P) + G + 1) / 2) * 2;
  static constexpr int kT = P / 2;                         // pair-stage iterations
  // pair-stage iterations per loop trip: 2 for the 16-lane groups (three chains per iteration; measured 6.46 ->
  // 6.15 ms at m = 40, d = 3), 1 for the 8-lane groups (four chains; unrolling costs 1 %): profiles/r02_variants.log
  static constexpr int kPairUnroll = (G == 16) ? 2 : 1;
  static constexpr int PX = NB * G;                        // coordinate row stride (>= P, even)
  static constexpr int kX = DD * PX;
  static constexpr int kNug = PX;
  static constexpr int kZ = PX;
  static constexpr int kI = PX / 2;                        // PX int32 compacted ids
  static constexpr int kRawI = PX / 2;                     // PX int32 raw ids of a row (as stored)
  static constexpr int kMeta = 2;                          // cond mask (8 B), row (4 B), pad
  static constexpr int kStage = kX + kNug + kZ + kI + kRawI + kMeta;
  static constexpr int kOffNug = kX, kOffZ = kX + kNug, kOffIds = kX + kNug + kZ,
                       kOffRaw = kX + kNug + kZ + kI, kOffMeta = kX + kNug + kZ + kI + kRawI;
  // resident warps the register allocation is sized for: three bands of a G = 8 group fit 168 registers (3 blocks
  // of 4 warps); everything else holds 255 registers, 8 warps per SM.  Those 8 warps are ONE block of 256 threads
  // (round 2; two blocks of 128 before): the per-block tables exist once per SM, which is what makes room, inside
  // the 227 KB of shared memory, for the conflict-free copy of the 2^(j/64) table (8 KB) and for the window of the
  // general-nu coefficient table (10 KB) next to the 32 sets in flight.
  static constexpr int kMinBlocks = (G == 8 && NB == 3) ? GPV_BAND_MINB3 : 1;
  static constexpr int kWPB = (kMinBlocks == 1) ? 8 : 4;   // warps per block
  static constexpr int kThreads = 32 * kWPB;
  static constexpr int kSetsPerWarp = 32 / G;
  static constexpr int kRaw = kBuf + 2 * kStage;
  // every set starts on a 128-byte line plus a skew of {0, 64, 32, 96} bytes: the row segments of the
  // sets of a half-warp tile one line, and the broadcast words of a warp instruction fall into
  // different banks
  static constexpr int kDoubles = ((kRaw + 15) / 16) * 16 + 16;
  static constexpr int kBytesPerBlock = kDoubles * 8 * kSetsPerWarp * kWPB;
  // replicated exp table / shared general-nu window: only where the 227 KB hold them next to the sets in flight
  // (P = 32 does not leave the room)
  static constexpr bool kWideTables = (kMinBlocks == 1) && (kBytesPerBlock + 13312 + 1024 <= 232448);
  static constexpr int kExpStride = kWideTables ? 16 : 1;
  static constexpr int kGenWin = kWideTables ? GPV_TAB_WIN_OCT * (1 << GPV_TAB_SUBBITS) : 0;   // intervals of the general-nu table kept in shared memory (cov_setup.h kTabWindow)
  static_assert(G == 8 || G == 16, "lane groups of 8 or 16");
  static_assert(P > 2 * G && P <= 4 * G, "band-folded kernel: 2G < P <= 4G");
  static_assert(G == 8 || NB == 3, "four bands of 16 lanes do not fit the register file");
};

template <int G> __host__ __device__ constexpr int band_of(int r) { return r / G; }
template <int G> __host__ __device__ constexpr int band_owner(int r) {   // lane (within the group) that holds row r
  return ((r / G) & 1) ? (G * (r / G) + G - 1 - r) : (r - G * (r / G));
}
template <int G> __host__ __device__ constexpr int band_row(int b, int q) {
  return (b & 1) ? ((b + 1) * G - 1 - q) : (b * G + q);
}

// Per (iteration t, lane q): where the pairs (i, i + t mod P) of the lane's points go in the packed
// staged triangle (byte offsets, 16 bits each; inactive combinations point at the lane's dump slot) and
// which staged point is the partner (index i + t mod P, 16 bits each).
template <int G, int P, int D>
__device__ __forceinline__ void build_store_table_band(uint4* __restrict__ stab) {
  using LY = BandLayout<G, P, D>;
  for (int idx = threadIdx.x; idx < LY::kT * G; idx += blockDim.x) {
    const int t = idx / G + 1, q = idx % G;
    const bool full = (2 * t < P);                       // even P: t = P/2 is covered by i < P/2 only
    unsigned off[4], par[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = band_row<G>(b, q);
      const int ic = i < P ? i : P - 1;
      int j = ic + t; if (j >= P) j -= P;
      par[b] = j;
      off[b] = LY::kScratch + q;                         // own dump slot: no write-write race
      if (b < LY::NB && i < P && (full || i < P / 2)) {
        const int a = i > j ? i : j, bb = i > j ? j : i;
        off[b] = tri_col(bb, P) + a - bb;
      }
      off[b] *= 8;
    }
    stab[idx] = make_uint4(off[0] | (off[1] << 16), off[2] | (off[3] << 16), par[0] | (par[1] << 16), par[2] | (par[3] << 16));
  }
}

// pair stage: point i evaluates the covariances (i, i + t mod P), t = 1..P/2; a lane carries its
// NB points through the same iteration (NB independent chains).
template <int KIND, int G, int P, int D, class C>
__device__ __forceinline__ void pair_stage_band_impl(const C& q, double* __restrict__ As,
                                                     const double* __restrict__ xs,
                                                     const double (&x)[BandLayout<G, P, D>::NB][BandLayout<G, P, D>::DD],
                                                     int gl, const uint4* __restrict__ stab,
                                                     const double* __restrict__ etab,
                                                     const unsigned* __restrict__ gtab, int d) {
  using LY = BandLayout<G, P, D>;
  constexpr int NB = LY::NB;
  constexpr int ES = (KIND == COV_GENERAL) ? 1 : LY::kExpStride;
  constexpr int GW = (KIND == COV_GENERAL) ? LY::kGenWin : 0;
  const uint4* stab_lane = stab + gl;
  char* Asb = reinterpret_cast<char*>(As);
  const double guard = (KIND == COV_GENERAL) ? 0.0 : kMathC[7];
  constexpr int UNR = (KIND == COV_GENERAL) ? 1 : LY::kPairUnroll;   // the general branch spills when unrolled
#pragma unroll UNR
  for (int t = 1; t <= LY::kT; ++t) {
    const uint4 offs = stab_lane[(t - 1) * G];
    const int jp[4] = {(int)(offs.z & 0xffffu), (int)(offs.z >> 16), (int)(offs.w & 0xffffu), (int)(offs.w >> 16)};
    const unsigned so[4] = {offs.x & 0xffffu, offs.x >> 16, offs.y & 0xffffu, offs.y >> 16};
    double r2[NB], v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) r2[b] = pair_r2<D>(xs, LY::PX, x[b], jp[b], d, guard);
    if constexpr (KIND == COV_GENERAL) {
      int idx[NB];
      bool sp = false, outside = false;
#pragma unroll
      for (int b = 0; b < NB; ++b) sp |= cov_general_special(r2[b], q.tab, &idx[b]);
      if constexpr (GW > 0) {
        const int win0 = q.tab.win0;
#pragma unroll
        for (int b = 0; b < NB; ++b) { idx[b] -= win0; outside |= (unsigned)idx[b] >= (unsigned)GW; }
      }
      if (__any_sync(0xffffffffu, sp)) {
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = cov_general_slow(r2[b], q, etab);
      } else if (GW > 0 && !__any_sync(0xffffffffu, outside)) {
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = cov_general_fast_shared<(GW > 0 ? GW : 1)>(r2[b], gtab + idx[b], gtab + (kTabDeg + 1) * GW + idx[b]);
      } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = cov_general_fast(r2[b], idx[b] + (GW > 0 ? q.tab.win0 : 0), q.tab);
      }
    } else {
      cov_eval_n<KIND, NB, C, ES>(r2, v, q, etab);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) *reinterpret_cast<double*>(Asb + so[b]) = v[b];
  }
}
// Closed forms: compiled as a function of its own.  Inlined into the 255-register kernel, ptxas
// schedules the chains one after the other on shared registers; as a separate function that takes
// the lane's coordinates by reference it interleaves them (measured: DESIGN.md 4.3).  The covariance
// constants go by value so that the kernel parameters are not forced into local memory.
template <int KIND, int G, int P, int D>
static __device__ __noinline__ void pair_stage_band_fn(CovConsts cc, double* __restrict__ As,
                                                       const double* __restrict__ xs,
                                                       const double (&x)[BandLayout<G, P, D>::NB][BandLayout<G, P, D>::DD],
                                                       int gl, const uint4* __restrict__ stab,
                                                       const double* __restrict__ etab, int d) {
  pair_stage_band_impl<KIND, G, P, D, CovConsts>(cc, As, xs, x, gl, stab, etab, nullptr, d);
}
template <int KIND, int G, int P, int D>
__device__ __forceinline__ void pair_stage_band(const UParams& q, double* __restrict__ As,
                                                const double* __restrict__ xs,
                                                const double (&x)[BandLayout<G, P, D>::NB][BandLayout<G, P, D>::DD],
                                                int gl, const uint4* __restrict__ stab,
                                                const double* __restrict__ etab,
                                                const unsigned* __restrict__ gtab, int d) {
  if constexpr (KIND != COV_GENERAL) {
    const CovConsts cc = {q.c0, q.c1, q.c2, q.c3, q.c4};
    pair_stage_band_fn<KIND, G, P, D>(cc, As, xs, x, gl, stab, etab, d);
  } else {
    pair_stage_band_impl<KIND, G, P, D, UParams>(q, As, xs, x, gl, stab, etab, gtab, d);
  }
}

template <int G, int P, int D, bool GENERAL>
__global__ void __launch_bounds__(BandLayout<G, P, D>::kThreads, BandLayout<G, P, D>::kMinBlocks)
u_band_kernel(const UParams q) {
  using LY = BandLayout<G, P, D>;
  constexpr int NB = LY::NB, SETS = LY::kSetsPerWarp, PX = LY::PX;
  constexpr int S0 = LY::S0, S1 = LY::S1, S2 = LY::S2, S3 = LY::S3;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned GMASK = (1u << G) - 1u;
  constexpr int kSelfLane = band_owner<G>(P - 1);    // lane that holds row P-1 (last band)

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
  const int d = (D > 0) ? D : q.d;
  const int p = q.p;

  double* buf = smem + (size_t)(warp * SETS + sub) * LY::kDoubles + ((sub & 1) * 8 + (sub >> 1) * 4);
  double* stage0 = buf + LY::kBuf;

  __shared__ uint4 stab[LY::kT * G];
  // 2^(j/64): 64 doubles, or (closed forms of the 256-thread kernels) every entry 16 times so that lane l reads
  // column l % 16 -- a half-warp's lookups then fall into 16 different bank pairs for any j
  constexpr int ES = GENERAL ? 1 : LY::kExpStride;
  __shared__ double etab_s[64 * ES];
  build_store_table_band<G, P, D>(stab);
  for (int i = threadIdx.x; i < 64 * ES; i += blockDim.x) etab_s[i] = kExp2Tab[i / ES];
  const double* etab = etab_s + (ES > 1 ? (lane & (ES - 1)) : 0);
  // general nu: a window of kGenWin intervals of the coefficient table (24 octaves of w, placed by the host where
  // the handle's neighbour distances are: CovTable::win0) in shared memory, coefficient-major like the global
  // table; pairs outside it take the global-memory path
  constexpr int GW = GENERAL ? LY::kGenWin : 0;
  // stored as a plane of high words followed by a plane of low words (bessel_table.cuh: cov_general_fast_shared)
  __shared__ unsigned gtab[(GW > 0) ? 2 * (kTabDeg + 1) * GW : 1];
  if constexpr (GW > 0) {
    const int win0 = q.tab.win0;
    for (int i = threadIdx.x; i < (kTabDeg + 1) * GW; i += blockDim.x) {
      const int k = i / GW, j = i % GW + win0;
      const double c = (j < q.tab.nint) ? q.tab.coef[tab_coef_index(k, j)] : 0.0;
      gtab[i] = (unsigned)__double2hiint(c);
      gtab[(kTabDeg + 1) * GW + i] = (unsigned)__double2loint(c);
    }
  }
  __syncthreads();

  // my rows; a row that does not exist (>= P, last band only) is clamped to P-1 and masked
  int rc[NB];
  bool vb[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int r = band_row<G>(b, gl);
    vb[b] = r < P;
    rc[b] = vb[b] ? r : P - 1;
  }

  double acc_quad = 0.0, acc_logd = 0.0, acc_qden = 0.0, acc_lden = 0.0;
  const int64_t stride = (int64_t)gridDim.x * LY::kWPB * SETS;
  const int64_t first = ((int64_t)blockIdx.x * LY::kWPB + warp) * SETS;

  // ---- input pipeline (cp.async; same two-stage scheme as u_sets_kernel) ---------------------------
  auto fetch_raw = [&](int64_t sidx, double* st) {
    int* raw = reinterpret_cast<int*>(st + LY::kOffRaw);
    double* meta = st + LY::kOffMeta;
    int* metai = reinterpret_cast<int*>(meta + 1);
    if (sidx < q.nsets) {
      const int32_t* nnr = q.nn + sidx * (int64_t)p;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const int e = gl + G * c;
        if (e < p) __pipeline_memcpy_async(raw + e, nnr + e, 4); else raw[e] = -1;
      }
      if (gl == 0) {
        __pipeline_memcpy_async(meta, q.cond + sidx, 8);
        if (q.rowmap != nullptr) __pipeline_memcpy_async(metai, q.rowmap + sidx, 4);
        else metai[0] = (int)(q.set_base + sidx);
      }
    } else {
#pragma unroll
      for (int c = 0; c < NB; ++c) raw[gl + G * c] = -1;
      if (gl == 0) { reinterpret_cast<unsigned long long*>(meta)[0] = 0ull; metai[0] = -1; }
    }
  };
  // compaction of the raw ids (U_NZentries.cpp:41-45; entries gl + G c of the row), then coordinates,
  // nuggets and z of my points -> stage st.  Returns n0.
  auto gather = [&](double* st) -> int {
    const int* raw = reinterpret_cast<const int*>(st + LY::kOffRaw);
    int* ids = reinterpret_cast<int*>(st + LY::kOffIds);
    double* xs = st;
    double* nug = st + LY::kOffNug;
    int rawv[NB];
    unsigned bal[NB];
    int n0 = 0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      rawv[c] = raw[gl + G * c];
      bal[c] = (__ballot_sync(FULL, rawv[c] >= 0) >> base) & GMASK;
      n0 += __popc(bal[c]);
    }
    const int npad = P - n0;
    const unsigned below = (1u << gl) - 1u;
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (gl + G * c < npad) ids[gl + G * c] = -1;
    __syncwarp();
    int pre = npad;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (rawv[c] >= 0) ids[pre + __popc(bal[c] & below)] = rawv[c];
      pre += __popc(bal[c]);
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
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

#if GPV_BAND_STAGGER > 0 && !defined(GPV_SIMT_EMU)
  // the two warps of a scheduler (w and w + kWPB/2) would otherwise run their phases in lock step -- both in the
  // fp64-bound pair stage, then both in the shared-memory-bound factorisation; start the second one half a batch
  // later so that the phases interleave (identical work per batch: the offset persists)
  if (LY::kWPB == 8 && warp >= LY::kWPB / 2) {
    const long long t0 = clock64();
    while (clock64() - t0 < GPV_BAND_STAGGER) {}
  }
#endif
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

    // ---- 1./2. my points of the current set ----------------------------------------------------------
    const double* xs = st;
    const int* ids = reinterpret_cast<const int*>(st + LY::kOffIds);
    const double* nugs = st + LY::kOffNug;
    double x[NB][LY::DD];
    double dg[NB];
    int id[NB];
    bool cd[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
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
      pair_stage_band<COV_GENERAL, G, P, D>(q, buf, xs, x, gl, stab, etab, gtab, d);
    } else {
      switch (q.cov) {
        case COV_EXP: pair_stage_band<COV_EXP, G, P, D>(q, buf, xs, x, gl, stab, etab, gtab, d); break;
        case COV_M15: pair_stage_band<COV_M15, G, P, D>(q, buf, xs, x, gl, stab, etab, gtab, d); break;
        case COV_M25: pair_stage_band<COV_M25, G, P, D>(q, buf, xs, x, gl, stab, etab, gtab, d); break;
        default: pair_stage_band<COV_ESQE, G, P, D>(q, buf, xs, x, gl, stab, etab, gtab, d); break;
      }
    }
    if (__any_sync(FULL, npad > 0)) {
      // padding occupies the leading indices: zero columns 0..npad-1 of the staged triangle
      __syncwarp();
      for (int j = 0; j < npad; ++j) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (vb[b] && rc[b] > j) buf[tri_col(j, P) + rc[b] - j] = 0.0;
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (vb[b]) buf[tri_col(rc[b], P)] = dg[b];
    __syncwarp();

#include "u_band_factor.inc"
#include "u_band_finish.inc"
    __pipeline_wait_prior(0);                              // set i+1 staged, ids of set i+2 landed
    __syncwarp();
    n0 = n0_next;
    bsel ^= 1;
  }

  // ---- deterministic block reduction of the likelihood partial sums ---------------------------------
  if (q.partials != nullptr) {
    __shared__ double red[LY::kWPB][4];
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
      for (int w = 0; w < LY::kWPB; ++w) a += red[w][threadIdx.x];
      q.partials[4 * blockIdx.x + threadIdx.x] = a;
    }
  }
}

}  // namespace gpv
