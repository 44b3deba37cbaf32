// tests/simt_emu/cuda_runtime.h -- TEST INFRASTRUCTURE.  A host stand-in for the CUDA builtins the set
// kernels use (gpvecchia_b200/csrc/u_kernels.cuh, u_band.cuh, bessel_table.cuh), so that the SAME kernel
// source can be compiled with g++ and run on the CPU, one host thread per CUDA thread:
//   * a block runs as blockDim.x std::threads; __syncthreads / __syncwarp are pthread barriers (per block /
//     per warp), so shared memory is only exchanged where the kernel synchronises -- a missing __syncwarp
//     shows up as a data race here (and under -fsanitize=thread) instead of passing by warp lockstep;
//   * __shfl_sync / __shfl_xor_sync / __ballot_sync / __any_sync exchange through a per-warp slot array
//     between two barrier waits (full-mask, convergent use only: that is all the kernels do);
//   * __pipeline_memcpy_async copies immediately (one legal execution of cp.async), commit / wait are no-ops;
//   * __shared__ is `static` (one block runs at a time); the dynamic shared array is provided by the harness;
//   * the MUFU seeds (rsqrt.approx / rcp.approx) are modelled as the exact value truncated to its high word.
// What this checks is the kernels' LOGIC (index maps, compaction, elimination order, outputs) against the
// oracle without a GPU; timing, register allocation and the real memory model are not modelled.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <cmath>
#include <cstdint>
#include <cstring>

#define GPV_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static const
#define __align__(n) alignas(n)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct double2 { double x, y; } __attribute__((aligned(16)));
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
typedef void* cudaStream_t;

namespace emu {
struct Warp {
  pthread_barrier_t bar;
  uint64_t slot[32];
};
struct Block {
  pthread_barrier_t bar;
  Warp* warps;
};
extern thread_local Block* tl_block;
extern thread_local Warp* tl_warp;
extern thread_local int tl_lane;
}  // namespace emu
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&emu::tl_block->bar); }
#ifdef GPV_EMU_DROP_SYNCWARP   // self-test of the race detection: without the kernel's __syncwarp calls TSAN must complain
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#else
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::tl_warp->bar); }
#endif
static inline uint64_t emu_exchange(uint64_t mine, int src) {
  emu::Warp* w = emu::tl_warp;
  w->slot[emu::tl_lane] = mine;
  pthread_barrier_wait(&w->bar);
  const uint64_t v = w->slot[src & 31];
  pthread_barrier_wait(&w->bar);
  return v;
}
static inline double __shfl_sync(unsigned, double v, int src) {
  uint64_t u; std::memcpy(&u, &v, 8); u = emu_exchange(u, src); double r; std::memcpy(&r, &u, 8); return r;
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)(uint32_t)emu_exchange((uint32_t)v, src); }
static inline double __shfl_xor_sync(unsigned m, double v, int mask) { return __shfl_sync(m, v, emu::tl_lane ^ mask); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emu::Warp* w = emu::tl_warp;
  w->slot[emu::tl_lane] = pred ? 1u : 0u;
  pthread_barrier_wait(&w->bar);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(w->slot[i] & 1u) << i;
  pthread_barrier_wait(&w->bar);
  return r;
}
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
static inline unsigned __activemask() { return 0xffffffffu; }

static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
using std::fma; using std::sqrt; using std::exp; using std::log; using std::fabs; using std::floor; using std::pow;
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline long long atomicMin(long long* p, long long v) {
  long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
// MUFU.RSQ64H / MUFU.RCP64H: about 20 good bits, low word zero
static inline double emu_rsqrt_seed(double a) { return __hiloint2double(__double2hiint(1.0 / std::sqrt(a)), 0); }
double* emu_dynamic_smem();   // harness: the block's dynamic shared memory (16-byte aligned)
static inline double emu_rcp_seed(double a) { return __hiloint2double(__double2hiint(1.0 / a), 0); }
