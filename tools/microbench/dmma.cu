// fp64 tensor-core MMA (mma.sync m8n8k4 f64) throughput / latency on sm_100a, alone and mixed with DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, int NF>   // CH independent DMMA accumulators, NF independent DFMA chains interleaved
__global__ void k(double* out, long long* cyc, int iters) {
  double c0[CH > 0 ? CH : 1], c1[CH > 0 ? CH : 1], f[NF > 0 ? NF : 1];
  for (int i = 0; i < CH; ++i) { c0[i] = i; c1[i] = -i; }
  for (int i = 0; i < NF; ++i) f[i] = 1.0 + i + threadIdx.x;
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = fma(f[i], 1.0000001, 1e-9);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
  long long h; const int it = 2048;
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); return (double)h / it; };
  k<1, 0><<<1, 32>>>(d, c, it); printf("DMMA dependent latency (1 warp): %.1f cycles\n", rd());
  k<8, 0><<<1, 32>>>(d, c, it); { double v = rd(); printf("DMMA 8 independent, 1 warp: %.1f cyc/iter = %.2f cyc per DMMA\n", v, v / 8); }
  for (int w : {4, 8, 16}) { k<8, 0><<<1, 32 * w>>>(d, c, it); double v = rd(); printf("DMMA 8 independent, %2d warps on one SM: %.1f cyc/iter -> %.2f SM-cycles per DMMA (= %.1f DFMA-equivalents/cycle/SM)\n", w, v, v / (8.0 * w), 256.0 * 8 * w / v / 32.0); }
  for (int w : {4, 16}) { k<0, 8><<<1, 32 * w>>>(d, c, it); double v = rd(); printf("DFMA only 8 chains, %2d warps: %.1f cyc/iter\n", w, v); }
  for (int w : {4, 16}) { k<8, 8><<<1, 32 * w>>>(d, c, it); double v = rd(); printf("DMMA x8 + DFMA x8 per iter, %2d warps: %.1f cyc/iter\n", w, v); }
  for (int w : {16}) { k<8, 32><<<1, 32 * w>>>(d, c, it); double v = rd(); printf("DMMA x8 + DFMA x32 per iter, %2d warps: %.1f cyc/iter\n", w, v); }
  for (int w : {16}) { k<0, 32><<<1, 32 * w>>>(d, c, it); double v = rd(); printf("DFMA x32 per iter, %2d warps: %.1f cyc/iter\n", w, v); }
  return 0;
}
