// u_band_ws.cuh -- EXPERIMENT (candidate 1; not the default path, not yet timed on a B200):
// a warp-specialised variant of the band-folded set kernel for 24 < P <= 32 (G = 8, four bands).
//
// Why: u_band_kernel holds 255 registers and 6.6 KB of shared memory per set, which leaves two warps per
// scheduler; the fp64 pipe is 47 % busy and what remains is latency inside a warp's dependent phases
// (profiles/r01_u_band_closed_P31_D2_nu15_phases.txt).  The pair stage needs few registers and has four
// independent chains per lane; the factorisation needs the registers and sits in step heads.  Here they run in
// different warps of one 384-thread block:
//   * warps 0-3 (one warp group, `setmaxnreg.inc` to 240 registers): CONSUMERS.  Each owns two batch slots of
//     four sets and runs steps 4-7 (u_band_factor.inc, the same text u_band_kernel includes) on a slot once
//     its four sets are marked full, then marks the slot empty.
//   * warps 4-11 (two warp groups, `setmaxnreg.dec` to 128): PRODUCERS, warp-per-set.  Producers 2c and 2c+1
//     fill sets {0,1} and {2,3} of consumer c's next slot: ids -> compaction (one ballot) -> coordinates,
//     nugget, z of point `lane` -> the lane's 15 or 16 pairs (lane, lane + t mod P), four at a time -> staged
//     triangle, diagonal, padding -> full.
// 12 warps per SM instead of 8 inside the same register file (4*32*240 + 8*32*128 = 63 488 of the 64 512 the launch holds) and shared memory
// (32 set buffers of triangle + one input stage = 176 KB).  Slots are handed over through monotone counters in
// shared memory (release store by one lane after __syncwarp, acquire spin by one lane before __syncwarp).
// Results are those of u_band_kernel up to the order of nothing: the same pairs, the same arithmetic per pair,
// the same factorisation text; checked on the host by tests/test_simt_emu.py.
#pragma once
#include "u_band.cuh"

namespace gpv {

constexpr int kWsConsumerWarps = 4;
constexpr int kWsProducerWarps = 8;
constexpr int kWsThreads = 32 * (kWsConsumerWarps + kWsProducerWarps);
constexpr int kWsSlots = 2;                      // batches in flight per consumer warp
constexpr int kWsConsumerRegs = 240, kWsProducerRegs = 128;
// GPV_WS_FINISH_IN_PRODUCERS = 1: the consumers stop after the factorisation (steps 4-5) and the producer warps
// run the sweep and the outputs (steps 6-7: u_band_finish.inc) as a second task, taking turns per batch; a
// consumer then only loads rows and eliminates.  Everything the sweep needs is in the slot: the unscaled columns,
// 1 / d_k in the diagonal slots, d_{P-1} in the last one, and the failure flag in the stage's raw-id words.
#ifndef GPV_WS_FINISH_IN_PRODUCERS
#define GPV_WS_FINISH_IN_PRODUCERS 0
#endif

template <int P, int D>
struct WsLayout {
  using LY = BandLayout<8, P, D>;
  static constexpr int kSet = ((LY::kBuf + LY::kStage + 15) / 16) * 16 + 16;   // triangle + ONE input stage (+ skew room)
  static constexpr int kSetBuffers = kWsConsumerWarps * kWsSlots * 4;
  static constexpr int kBytesPerBlock = kSetBuffers * kSet * 8;
  static constexpr int kSetsPerBlock = kWsConsumerWarps * 4;                     // per pass of the persistent loop
  static_assert(P > 24 && P <= 32, "warp-specialised variant: four bands of eight lanes");
};

// ---- slot hand-over --------------------------------------------------------------------------------------
__device__ __forceinline__ void ws_signal(int* flag, int value) {   // one lane, after the warp's __syncwarp
#ifdef GPV_SIMT_EMU
  __atomic_store_n(flag, value, __ATOMIC_RELEASE);
#else
  __threadfence_block();
  *reinterpret_cast<volatile int*>(flag) = value;
#endif
}
__device__ __forceinline__ void ws_wait(const int* flag, int value) {   // one lane; the warp's __syncwarp follows
#ifdef GPV_SIMT_EMU
  while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) < value) sched_yield();
#else
  while (*reinterpret_cast<const volatile int*>(flag) < value) __nanosleep(40);
  __threadfence_block();
#endif
}
template <int REGS> __device__ __forceinline__ void ws_regs_inc() {
#ifndef GPV_SIMT_EMU
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}
template <int REGS> __device__ __forceinline__ void ws_regs_dec() {
#ifndef GPV_SIMT_EMU
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}

// ---- producer: one set by one warp -------------------------------------------------------------------------
// Pairs of point r = lane: (r, r + t mod P), t = 1..P/2 (for even P the last t only for r < P/2): every unordered
// pair once, 15 or 16 per lane.  Four values of t per trip give the lane four independent covariance chains.
// Where pair (r, t) goes in the packed staged triangle and who the partner is depends only on (r, t): a per-block
// table holds the offset (in doubles; 0xffff: inactive, the store goes to the lane's dump word) and the partner.
constexpr int kWsTrips = 4;                       // trips of four pairs: t = 1..16 covers P/2 <= 16
template <int P>
__device__ __forceinline__ void ws_build_pair_table(unsigned* __restrict__ wtab) {
  constexpr int T = P / 2;
  for (int idx = threadIdx.x; idx < kWsTrips * 4 * 32; idx += blockDim.x) {
    const int t = idx / 32 + 1, r = idx % 32;
    const int rr = r < P ? r : P - 1;
    const int tt = t <= T ? t : T;
    int j = rr + tt; if (j >= P) j -= P;
    const bool act = (r < P) && (t <= T) && (2 * t < P || r < P / 2);
    const int a = rr > j ? rr : j, b = rr > j ? j : rr;
    wtab[idx] = (act ? (unsigned)(tri_col(b, P) + a - b) : 0xffffu) | ((unsigned)j << 16);
  }
}
template <int KIND, int P, int D>
static __device__ __noinline__ void ws_pair_stage(CovConsts cc, double* __restrict__ buf, const double* __restrict__ xs,
                                                  const double (&x)[BandLayout<8, P, D>::DD], int lane,
                                                  const unsigned* __restrict__ wtab, double* __restrict__ dump,
                                                  const double* __restrict__ etab, int d) {
  using LY = BandLayout<8, P, D>;
  const double guard = kMathC[7];
  const unsigned* wl = wtab + lane;
#pragma unroll 1
  for (int trip = 0; trip < kWsTrips; ++trip) {
    double r2[4], v[4];
    unsigned ent[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ent[u] = wl[(trip * 4 + u) * 32];
      r2[u] = pair_r2<D>(xs, LY::PX, x, (int)(ent[u] >> 16), d, guard);
    }
    cov_eval_n<KIND, 4>(r2, v, cc, etab);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned off = ent[u] & 0xffffu;
      double* dst = (off != 0xffffu) ? buf + off : dump;
      *dst = v[u];
    }
  }
}

// id of entry `lane` of set sidx as stored (-1: missing, past the row, or past the last set)
__device__ __forceinline__ int ws_load_raw(const UParams& q, int64_t sidx, int lane, int p) {
  return (sidx < q.nsets && lane < p) ? q.nn[sidx * (int64_t)p + lane] : -1;
}

// General-nu table path (bessel_table.cuh): one warp vote per trip decides between the table and the per-lane
// slow path (zero distance, far pairs, arguments outside the table), as in u_band_kernel.  Inlined: the table
// descriptor lives in the kernel parameters.
template <int P, int D>
__device__ __forceinline__ void ws_pair_stage_general(const UParams& q, double* __restrict__ buf,
                                                      const double* __restrict__ xs,
                                                      const double (&x)[BandLayout<8, P, D>::DD], int lane,
                                                      const unsigned* __restrict__ wtab, double* __restrict__ dump,
                                                      const double* __restrict__ etab, int d) {
  using LY = BandLayout<8, P, D>;
  const unsigned* wl = wtab + lane;
#pragma unroll 1
  for (int trip = 0; trip < kWsTrips; ++trip) {
    double r2[4], v[4];
    unsigned ent[4];
    int idx[4];
    bool sp = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ent[u] = wl[(trip * 4 + u) * 32];
      r2[u] = pair_r2<D>(xs, LY::PX, x, (int)(ent[u] >> 16), d, 0.0);
      sp |= cov_general_special(r2[u], q.tab, &idx[u]);
    }
    if (__any_sync(0xffffffffu, sp)) {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = cov_general_slow(r2[u], q, etab);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = cov_general_fast(r2[u], idx[u], q.tab);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned off = ent[u] & 0xffffu;
      double* dst = (off != 0xffffu) ? buf + off : dump;
      *dst = v[u];
    }
  }
}

// What a producer lane knows about point `lane` of a set before the set's buffer is free: gathered one set ahead,
// so the loads travel while the previous set's pair stage runs.
template <int DD>
struct WsPoint {
  int id, n0, rowv;
  unsigned long long cmask;     // lane 0 only
  double x[DD], nv, zv;
};

// ids of the row as stored -> compaction like inds.elem(find(inds)) (U_NZentries.cpp:41-45; missing entries become
// LEADING padding) through the warp's scratch words -> coordinates, nugget and z of point `lane` (loads issued
// here, consumed by ws_fill_set)
template <int P, int D>
__device__ __forceinline__ WsPoint<BandLayout<8, P, D>::DD> ws_gather_set(const UParams& q, int64_t sidx, int raw,
                                                                           int* __restrict__ scratch, int lane, int d) {
  using LY = BandLayout<8, P, D>;
  constexpr unsigned FULL = 0xffffffffu;
  WsPoint<LY::DD> pt;
  const bool live = sidx < q.nsets;
  const unsigned bal = __ballot_sync(FULL, raw >= 0);
  pt.n0 = __popc(bal);
  const int npad = P - pt.n0;
  __syncwarp();                                   // the previous set's reads of the scratch words are done
  scratch[lane] = -1;
  __syncwarp();
  if (raw >= 0) scratch[npad + __popc(bal & ((1u << lane) - 1u))] = raw;
  __syncwarp();
  pt.id = (lane < P) ? scratch[lane] : -1;
  pt.cmask = 0ull;
  pt.rowv = -1;
  if (lane == 0 && live) {
    pt.cmask = (unsigned long long)q.cond[sidx];
    pt.rowv = (q.rowmap != nullptr) ? q.rowmap[sidx] : (int)(q.set_base + sidx);
  }
#pragma unroll
  for (int c = 0; c < LY::DD; ++c) pt.x[c] = 0.0;
  pt.nv = 0.0; pt.zv = 0.0;
  if (pt.id >= 0) {
#pragma unroll
    for (int c = 0; c < LY::DD; ++c)
      if (c < d) pt.x[c] = q.locs[(int64_t)pt.id * d + c];
    pt.nv = q.nuggets[pt.id];
    if (q.zloc != nullptr) pt.zv = q.zloc[pt.id];
  }
  return pt;
}

// input stage, staged triangle, diagonal and padding of one set from its gathered points
template <int P, int D, bool GENERAL>
__device__ __forceinline__ void ws_fill_set(const UParams& q, const WsPoint<BandLayout<8, P, D>::DD>& pt,
                                            double* __restrict__ buf, double* __restrict__ st,
                                            const unsigned* __restrict__ wtab, double* __restrict__ dump,
                                            const double* __restrict__ etab, int lane, int d, int p) {
  using LY = BandLayout<8, P, D>;
  int* ids = reinterpret_cast<int*>(st + LY::kOffIds);
  double* xs = st;
  double* nug = st + LY::kOffNug;
  double* zs = st + LY::kOffZ;
  double* meta = st + LY::kOffMeta;
  int* metai = reinterpret_cast<int*>(meta + 1);
  const int npad = P - pt.n0;
  if (lane == 0) {
    reinterpret_cast<unsigned long long*>(meta)[0] = pt.cmask;
    metai[0] = pt.rowv;
    metai[1] = pt.n0;
  }
  const int r = lane;                 // point r = lane of the compacted set
  double x[LY::DD];
#pragma unroll
  for (int c = 0; c < LY::DD; ++c) x[c] = pt.x[c];
  if (r < P) {
    ids[r] = pt.id;
    if (D == 2) {
      reinterpret_cast<double2*>(xs)[r] = make_double2(x[0], x[1]);
    } else {
      for (int c = 0; c < d; ++c) xs[c * LY::PX + r] = x[c];
    }
    nug[r] = pt.nv;
    zs[r] = pt.zv;
  }
  __syncwarp();
  const unsigned long long cmask = reinterpret_cast<const unsigned long long*>(meta)[0];
  double dg = 1.0;
  if (r < P && pt.id >= 0) {
    // compacted entry j reads revCond[row, p - n0 + j] (:47); local index = npad + j
    const bool cd = (cmask >> ((r - (P - p)) & 63)) & 1ull;
    dg = q.c0 + clamp_nugget(pt.nv * (1.0 - (cd ? 1.0 : 0.0)));   // Inf * 0 = NaN kept
  }
  if (GENERAL) {
    ws_pair_stage_general<P, D>(q, buf, xs, x, lane, wtab, dump, etab, d);
  } else {
    const CovConsts cc = {q.c0, q.c1, q.c2, q.c3, q.c4};
    switch (q.cov) {
      case COV_EXP: ws_pair_stage<COV_EXP, P, D>(cc, buf, xs, x, lane, wtab, dump, etab, d); break;
      case COV_M15: ws_pair_stage<COV_M15, P, D>(cc, buf, xs, x, lane, wtab, dump, etab, d); break;
      case COV_M25: ws_pair_stage<COV_M25, P, D>(cc, buf, xs, x, lane, wtab, dump, etab, d); break;
      default: ws_pair_stage<COV_ESQE, P, D>(cc, buf, xs, x, lane, wtab, dump, etab, d); break;
    }
  }
  if (npad > 0) {
    // padding occupies the leading indices: zero columns 0..npad-1 of the staged triangle (only the first m rows
    // of a data set have padding, so this is rare; npad is warp-uniform)
    __syncwarp();
    for (int j = 0; j < npad; ++j)
      if (r < P && r > j) buf[tri_col(j, P) + r - j] = 0.0;
  }
  if (r < P) buf[tri_col(r, P)] = dg;
}

template <int P, int D, bool GENERAL>
__global__ void __launch_bounds__(kWsThreads, 1)
u_band_ws_kernel(const UParams q) {
  constexpr int G = 8;
  using LY = BandLayout<G, P, D>;
  using WL = WsLayout<P, D>;
  constexpr int NB = LY::NB;
  constexpr int S0 = LY::S0, S1 = LY::S1, S2 = LY::S2, S3 = LY::S3;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned GMASK = (1u << G) - 1u;
  constexpr int kSelfLane = band_owner<G>(P - 1);

#ifdef GPV_SIMT_EMU
  double* smem = emu_dynamic_smem();
#else
  extern __shared__ __align__(16) double smem[];
#endif
  __shared__ double etab[64];
  __shared__ int full_cnt[kWsConsumerWarps][kWsSlots][4];   // how many times set `sub` of the slot has been filled
  __shared__ int empty_cnt[kWsConsumerWarps][kWsSlots];     // how many times the slot has been consumed
  __shared__ int fact_cnt[kWsConsumerWarps][kWsSlots];      // how many times the slot has been factored (finish in producers)
  __shared__ double red[kWsConsumerWarps + kWsProducerWarps][4];
  __shared__ int pscratch[kWsProducerWarps][32];            // compaction scratch of each producer warp
  __shared__ double pdump[kWsProducerWarps][32];            // where a lane's inactive pair slots are stored
  __shared__ unsigned wtab[kWsTrips * 4 * 32];              // pair table of the producers (ws_build_pair_table)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int d = (D > 0) ? D : q.d;
  const int p = q.p;
  if (threadIdx.x < 64) etab[threadIdx.x] = kExp2Tab[threadIdx.x];
  ws_build_pair_table<P>(wtab);
  if (threadIdx.x < kWsConsumerWarps * kWsSlots * 4) (&full_cnt[0][0][0])[threadIdx.x] = 0;
  if (threadIdx.x < kWsConsumerWarps * kWsSlots) { (&empty_cnt[0][0])[threadIdx.x] = 0; (&fact_cnt[0][0])[threadIdx.x] = 0; }
  __syncthreads();

  // set buffer (consumer c, slot s, set sub): triangle then its input stage; the skew keeps the broadcast words
  // of the four sets of a consumer instruction in different banks (as in u_band_kernel)
  auto set_buf = [&](int c, int s, int sub) -> double* {
    return smem + (size_t)((c * kWsSlots + s) * 4 + sub) * WL::kSet + ((sub & 1) * 8 + (sub >> 1) * 4);
  };
  const int64_t stride = (int64_t)gridDim.x * WL::kSetsPerBlock;

  double acc_quad = 0.0, acc_logd = 0.0, acc_qden = 0.0, acc_lden = 0.0;
  if (warp < kWsConsumerWarps) {
    // ================================ consumer ================================
    ws_regs_inc<kWsConsumerRegs>();
    const int c = warp;
    const int sub = lane / G;
    const int gl = lane % G;
    const int base = sub * G;
    int rc[NB];
    bool vb[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int r = band_row<G>(b, gl);
      vb[b] = r < P;
      rc[b] = vb[b] ? r : P - 1;
    }
    const int64_t first = ((int64_t)blockIdx.x * kWsConsumerWarps + c) * 4;
    int k = 0;
    for (int64_t s0 = first; s0 < q.nsets; s0 += stride, ++k) {
      const int slot = k & 1;
      const int fill = (k >> 1) + 1;
      if (lane < 4) ws_wait(&full_cnt[c][slot][lane], fill);
      __syncwarp();
      double* buf = set_buf(c, slot, sub);
      double* st = buf + LY::kBuf;
      const uint64_t cmask = reinterpret_cast<const unsigned long long*>(st + LY::kOffMeta)[0];
      const int row = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[0];
      const int n0 = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[1];
      const bool row_ok = row >= 0;
      const int npad = P - n0;
      const int* ids = reinterpret_cast<const int*>(st + LY::kOffIds);
      const double* nugs = st + LY::kOffNug;
      int id[NB];
      bool cd[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        id[b] = vb[b] ? ids[rc[b]] : -1;
        cd[b] = (id[b] >= 0) ? (bool)((cmask >> ((rc[b] - (P - p)) & 63)) & 1ull) : false;
      }
#include "u_band_factor.inc"
#if GPV_WS_FINISH_IN_PRODUCERS
      (void)dlast; (void)nugs; (void)row_ok; (void)npad; (void)id; (void)cd;
      if (gl == 0) reinterpret_cast<int*>(st + LY::kOffRaw)[0] = fail ? 1 : 0;
      __syncwarp();
      if (lane == 0) ws_signal(&fact_cnt[c][slot], fill);
#else
#include "u_band_finish.inc"
      __syncwarp();
      if (lane == 0) ws_signal(&empty_cnt[c][slot], fill);
#endif
    }
  } else {
    // ================================ producer ================================
    ws_regs_dec<kWsProducerRegs>();
    const int pw = warp - kWsConsumerWarps;
    const int c = pw >> 1;
    const int h = pw & 1;
    const int64_t first = ((int64_t)blockIdx.x * kWsConsumerWarps + c) * 4;
#if GPV_WS_FINISH_IN_PRODUCERS
    // steps 6-7 of batch kb of consumer c, with the consumer's lane mapping (eight lanes per set, four sets)
    const int sub_f = lane / G, gl = lane % G, base = sub_f * G;
    int rc[NB];
    bool vb[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int r = band_row<G>(b, gl);
      vb[b] = r < P;
      rc[b] = vb[b] ? r : P - 1;
    }
    auto finish_batch = [&](int kb) {
      const int slot = kb & 1;
      const int fill = (kb >> 1) + 1;
      if (lane == 0) ws_wait(&fact_cnt[c][slot], fill);
      __syncwarp();
      double* buf = set_buf(c, slot, sub_f);
      double* st = buf + LY::kBuf;
      const uint64_t cmask = reinterpret_cast<const unsigned long long*>(st + LY::kOffMeta)[0];
      const int row = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[0];
      const int n0 = reinterpret_cast<const int*>(st + LY::kOffMeta + 1)[1];
      const bool row_ok = row >= 0;
      const int npad = P - n0;
      const int* ids = reinterpret_cast<const int*>(st + LY::kOffIds);
      const double* nugs = st + LY::kOffNug;
      int id[NB];
      bool cd[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        id[b] = vb[b] ? ids[rc[b]] : -1;
        cd[b] = (id[b] >= 0) ? (bool)((cmask >> ((rc[b] - (P - p)) & 63)) & 1ull) : false;
      }
      const bool fail = reinterpret_cast<const int*>(st + LY::kOffRaw)[0] != 0;
      const double dlast = buf[tri_col(P - 1, P)];
#include "u_band_finish.inc"
      __syncwarp();
      if (lane == 0) ws_signal(&empty_cnt[c][slot], fill);
    };
#endif
    // The producer's sets in order: j = 2 k + e is set 2 h + e of the consumer's batch k.  One set ahead of the one
    // being filled, the next set's points are gathered (loads in flight during the pair stage); two ahead, its ids.
    auto sidx_of = [&](int j) -> int64_t { return first + (int64_t)(j >> 1) * stride + 2 * h + (j & 1); };
    int* scratch = &pscratch[pw][0];
    int raw = ws_load_raw(q, sidx_of(1), lane, p);
    WsPoint<LY::DD> pt = ws_gather_set<P, D>(q, sidx_of(0), ws_load_raw(q, sidx_of(0), lane, p), scratch, lane, d);
    int k = 0;
    for (int64_t s0 = first; s0 < q.nsets; s0 += stride, ++k) {
      const int slot = k & 1;
      const int fill = (k >> 1) + 1;
      if (k >= kWsSlots) {                       // the previous use of this slot has been consumed
        if (lane == 0) ws_wait(&empty_cnt[c][slot], fill - 1);
        __syncwarp();
      }
#pragma unroll 1
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * k + e;
        const int sub = 2 * h + e;
        double* buf = set_buf(c, slot, sub);
        const WsPoint<LY::DD> cur = pt;
        pt = ws_gather_set<P, D>(q, sidx_of(j + 1), raw, scratch, lane, d);
        raw = ws_load_raw(q, sidx_of(j + 2), lane, p);
        ws_fill_set<P, D, GENERAL>(q, cur, buf, buf + LY::kBuf, wtab, &pdump[pw][lane], etab, lane, d, p);
        __syncwarp();
        if (lane == 0) ws_signal(&full_cnt[c][slot][sub], fill);
      }
#if GPV_WS_FINISH_IN_PRODUCERS
      if (k >= 1 && ((k - 1) & 1) == h) finish_batch(k - 1);
#endif
    }
#if GPV_WS_FINISH_IN_PRODUCERS
    if (k >= 1 && ((k - 1) & 1) == h) finish_batch(k - 1);       // the consumer's last batch
#endif
  }

  // ---- deterministic block reduction of the likelihood partial sums (held by the warps that run step 7) ------
  if (q.partials != nullptr) {
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
      for (int w = 0; w < kWsConsumerWarps + kWsProducerWarps; ++w) a += red[w][threadIdx.x];
      q.partials[4 * blockIdx.x + threadIdx.x] = a;
    }
  }
}

}  // namespace gpv
