// Shared-memory broadcast-load throughput on sm_100a (development aid; DESIGN.md quotes the result).
#include <cstdio>
#include <cuda_runtime.h>
template <int W, int GROUPS>   // W = bytes per lane (8/16), GROUPS = distinct addresses per warp
__global__ void lds_tp(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = lane / (32 / GROUPS);
  // group g reads from a region offset by (128/GROUPS) bytes mod 128 -> distinct banks
  const double* base = sm + g * (16 / GROUPS) + g * 512 + (threadIdx.x >> 5) * 64;
  double acc0 = 0, acc1 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (W == 16) {
        double2 v = *reinterpret_cast<const double2*>(base + ((i + u * 2) & 30));
        acc0 += v.x; acc1 += v.y;
      } else {
        acc0 += base[(i + u) & 31];
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_tp(double* out, long long* cyc, int iters) {
  double a = threadIdx.x, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += __shfl_sync(0xffffffffu, a + u, (i + u) & 31);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
  long long h; const int it = 2048;
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); return (double)h / (it * 8.0); };
  const int threads = 512;   // 16 warps on one SM
#define RUN(W, G) lds_tp<W, G><<<1, threads>>>(d, c, it); printf("LDS.%d, %d distinct addr/warp, 16 warps: %.2f cyc per warp-instr per SM (/16 warps = %.2f wavefront-cycles/instr)\n", W * 8, G, rd(), rd() / 16.0);
  RUN(8, 1) RUN(8, 2) RUN(8, 4) RUN(16, 1) RUN(16, 2) RUN(16, 4)
  shfl_tp<<<1, threads>>>(d, c, it); printf("SHFL f64 (2 SHFL.32), 16 warps: %.2f cyc per f64 shuffle per SM-slot (/16 = %.2f)\n", rd(), rd() / 16.0);
  return 0;
}
