// tests/simt_emu/emu_harness.cpp -- TEST INFRASTRUCTURE: runs the band-folded set kernel
// (gpvecchia_b200/csrc/u_band.cuh, unchanged source) on the CPU, one std::thread per CUDA thread, through
// the stand-in builtins of cuda_runtime.h in this directory.  Built and driven by tests/test_simt_emu.py:
//   g++ -std=c++17 -O1 -pthread -shared -fPIC -Itests/simt_emu -Igpvecchia_b200/csrc -Iinclude \
//       tests/simt_emu/emu_harness.cpp -o libemu.so
#include "cuda_runtime.h"   // the stand-in (this directory comes first on the include path)
#define GPV_DEFINE_TABLE_BUILDER
#include "bessel_table.cuh"
#include "cov_setup.h"
#include "u_band.cuh"

#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

thread_local emu::Block* emu::tl_block = nullptr;
thread_local emu::Warp* emu::tl_warp = nullptr;
thread_local int emu::tl_lane = 0;
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

static double* g_dyn_smem = nullptr;
static int g_family = 1;
double* emu_dynamic_smem() { return g_dyn_smem; }

namespace {

typedef void (*KernelFn)(const gpv::UParams);

void run_grid_fn(const std::function<void()>& kernel, int grid, int threads, size_t smem_bytes);
void run_grid(KernelFn kernel, int grid, int threads, size_t smem_bytes, const gpv::UParams& q) {
  run_grid_fn([&]() { kernel(q); }, grid, threads, smem_bytes);
}
void run_grid_fn(const std::function<void()>& kernel, int grid, int threads, size_t smem_bytes) {
  const int nwarps = threads / 32;
  for (int b = 0; b < grid; ++b) {
    void* mem = nullptr;
    if (posix_memalign(&mem, 128, smem_bytes + 128) != 0) std::abort();
    std::memset(mem, 0xFF, smem_bytes + 128);          // uninitialised shared memory reads as NaN
    g_dyn_smem = static_cast<double*>(mem);
    std::vector<emu::Warp> warps(nwarps > 0 ? nwarps : 1);
    emu::Block blk;
    blk.warps = warps.data();
    pthread_barrier_init(&blk.bar, nullptr, threads);
    for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t) {
      pool.emplace_back([&, t]() {
        emu::tl_block = &blk;
        emu::tl_warp = &warps[t / 32];
        emu::tl_lane = t % 32;
        threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b;
        blockDim.x = (unsigned)threads; gridDim.x = (unsigned)grid;
        kernel();
      });
    }
    for (auto& th : pool) th.join();
    for (auto& w : warps) pthread_barrier_destroy(&w.bar);
    pthread_barrier_destroy(&blk.bar);
    std::free(mem);
  }
  g_dyn_smem = nullptr;
}

template <int G, int P, int D>
void run_sets(int grid, const gpv::UParams& q) {
  run_grid(gpv::u_sets_kernel<G, P, D, false>, grid, gpv::kThreadsPerBlock, gpv::SetLayout<G, P, D>::kBytesPerBlock, q);
}
template <int G, int P, int D>
void run_band_general(int grid, const gpv::UParams& q) {
  run_grid(gpv::u_band_kernel<G, P, D, true>, grid, gpv::BandLayout<G, P, D>::kThreads, gpv::BandLayout<G, P, D>::kBytesPerBlock, q);
}
template <int G, int P, int D>
void run_band(int grid, const gpv::UParams& q) {
  run_grid(gpv::u_band_kernel<G, P, D, false>, grid, gpv::BandLayout<G, P, D>::kThreads, gpv::BandLayout<G, P, D>::kBytesPerBlock, q);
}

}  // namespace

// One launch of u_band_kernel<G, P, D, closed forms> on `grid` blocks.  Arrays as the C-ABI library lays them
// out on the device: locs [N][d] row-major, nn [nsets][p] int32 0-based with -1 = missing, cond one
// 64-bit mask per row, nuggets [N], out [nsets][p] row-major (row_off == NULL) or packed, zloc [N] or NULL,
// partials [grid][4] or NULL.  cov: 0 exp, 1 Matern 1.5, 2 Matern 2.5, 3 esqe; c[5] = c0..c4 as setup_cov
// (gpv_capi.cu) computes them.  Returns 0, or 1 if that instantiation is not compiled here.
// family 1 = u_band_kernel (three or four rows per lane), family 0 = u_sets_kernel (two rows per lane; D = 0 is
// the run-time-dimension instantiation).
static int emu_launch(int family, int G, int P, int D, int grid, const gpv::UParams& q) {
  if (family == 1) {
    if (G == 8 && P == 31 && D == 2) run_band<8, 31, 2>(grid, q);
    else if (G == 8 && P == 21 && D == 3) run_band<8, 21, 3>(grid, q);
    else if (G == 16 && P == 41 && D == 3) run_band<16, 41, 3>(grid, q);
    else return 1;
  } else {
    if (G == 16 && P == 31 && D == 2) run_sets<16, 31, 2>(grid, q);
    else if (G == 4 && P == 8 && D == 0) run_sets<4, 8, 0>(grid, q);
    else if (G == 8 && P == 11 && D == 3) run_sets<8, 11, 3>(grid, q);
    else if (G == 32 && P == 51 && D == 2) run_sets<32, 51, 2>(grid, q);
    else return 1;
  }
  return 0;
}
extern "C" int emu_u_band(int G, int P, int D, int grid, int64_t nsets, int p, int d, const double* locs,
                          const int32_t* nn, const uint64_t* cond, const double* nuggets, double* out,
                          const int64_t* row_off, const double* zloc, int full_z, double* partials,
                          unsigned long long* nfail, long long* first_fail, int cov, const double* c) {
  gpv::UParams q;
  std::memset(&q, 0, sizeof(q));
  q.nrows = nsets; q.nsets = nsets; q.set_base = 0; q.rowmap = nullptr; q.row0 = 0; q.p = p; q.d = d;
  q.locs = locs; q.nn = nn; q.cond = cond; q.nuggets = nuggets; q.out = out; q.row_off = row_off;
  q.zloc = zloc; q.full_z = full_z; q.skip_rows = 0; q.partials = partials; q.nfail = nfail;
  q.first_fail = first_fail; q.cov = cov;
  q.c0 = c[0]; q.c1 = c[1]; q.c2 = c[2]; q.c3 = c[3]; q.c4 = c[4];
  return emu_launch(g_family, G, P, D, grid, q);
}
extern "C" int emu_u_sets(int family, int G, int P, int D, int grid, int64_t nsets, int p, int d, const double* locs,
                          const int32_t* nn, const uint64_t* cond, const double* nuggets, double* out,
                          const int64_t* row_off, const double* zloc, int full_z, double* partials,
                          unsigned long long* nfail, long long* first_fail, int cov, const double* c) {
  g_family = family;
  const int rc = emu_u_band(G, P, D, grid, nsets, p, d, locs, nn, cond, nuggets, out, row_off, zloc, full_z, partials,
                            nfail, first_fail, cov, c);
  g_family = 1;
  return rc;
}

// General-nu Matern (Matern.cpp:72-83) through the table path: the coefficient table is built by the library's
// own build_cov_table_kernel (one 32-thread block per interval, emulated like the set kernels) with the host
// set-up of cov_setup.h, then u_band_kernel<G, P, D, general> runs on it.  w_max: squared bounding-box diagonal.
extern "C" int emu_u_band_general(int family, int G, int P, int D, int grid, int64_t nsets, int p, int d, const double* locs,
                                  const int32_t* nn, const uint64_t* cond, const double* nuggets, double* out,
                                  unsigned long long* nfail, long long* first_fail, double sig2, double range,
                                  double nu, double w_max) {
  gpv::UParams q;
  std::memset(&q, 0, sizeof(q));
  q.nrows = nsets; q.nsets = nsets; q.p = p; q.d = d;
  q.locs = locs; q.nn = nn; q.cond = cond; q.nuggets = nuggets; q.out = out;
  q.nfail = nfail; q.first_fail = first_fail;
  q.cov = gpv::COV_GENERAL; q.c0 = sig2; q.inv_range = 1.0 / range;
  gpv::nu_constants(nu, sig2, &q.tab);
  gpv::general_table_range(range, w_max, &q.tab);
  std::vector<double> coef((size_t)(gpv::kTabDeg + 1) * gpv::kTabStride, 0.0);
  q.tab.coef = coef.data();
  const gpv::CovTable t = q.tab;
  const double inv_range = q.inv_range;
  double* cp = coef.data();
  run_grid_fn([&]() { gpv::build_cov_table_kernel(t, inv_range, cp); }, t.nint, 32, 0);
  if (family == 1 && G == 8 && P == 31 && D == 2) run_band_general<8, 31, 2>(grid, q);
  else if (family == 1 && G == 16 && P == 41 && D == 3) run_band_general<16, 41, 3>(grid, q);
  else return 1;
  return 0;
}
