// Latency / throughput micro-benchmarks of the fp64 pipe and friends on sm_100a (development aid;
// numbers are quoted in DESIGN.md).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
  double a[CHAINS];
  for (int c = 0; c < CHAINS; ++c) a[c] = 1.0 + threadIdx.x + c;
  const double b = 1.0000001, d = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], b, d);
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void mufu_chain(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    a = y + 1.5;   // 1 DADD + MUFU per iteration
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = (double)((threadIdx.x + 1) & 31);
  __syncthreads();
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) idx = (int)sm[idx];
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
// seed accuracy of rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64
__global__ void seed_err(double* out) {
  double worst_rs = 0, worst_rc = 0;
  for (int i = 0; i < 200000; ++i) {
    double a = 1.0 + (threadIdx.x * 200000.0 + i) / (32 * 200000.0) * 3.0;
    double y, z;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(a));
    double e1 = fabs(y * sqrt(a) - 1.0), e2 = fabs(z * a - 1.0);
    worst_rs = e1 > worst_rs ? e1 : worst_rs;
    worst_rc = e2 > worst_rc ? e2 : worst_rc;
  }
  out[2 * threadIdx.x] = worst_rs;
  out[2 * threadIdx.x + 1] = worst_rc;
}

int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
  long long h; const int it = 4096;
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); return (double)h / it; };
  dfma_chain<1><<<1, 32>>>(d, c, it); printf("DFMA dependent latency (1 warp, 1 chain): %.2f cyc/iter\n", rd());
  dfma_chain<2><<<1, 32>>>(d, c, it); printf("DFMA 2 chains: %.2f cyc/iter (%.2f per DFMA)\n", rd(), rd() / 2);
  dfma_chain<4><<<1, 32>>>(d, c, it); printf("DFMA 4 chains: %.2f cyc/iter (%.2f per DFMA)\n", rd(), rd() / 4);
  dfma_chain<8><<<1, 32>>>(d, c, it); printf("DFMA 8 chains: %.2f cyc/iter (%.2f per DFMA)\n", rd(), rd() / 8);
  for (int w : {1, 2, 4, 8, 16}) { dfma_chain<1><<<1, 32 * w>>>(d, c, it); printf("DFMA 1 chain, %2d warps/block on one SM: %.2f cyc/iter\n", w, rd()); }
  for (int w : {4, 8, 16}) { dfma_chain<2><<<1, 32 * w>>>(d, c, it); printf("DFMA 2 chains, %2d warps: %.2f cyc/iter\n", w, rd()); }
  mufu_chain<<<1, 32>>>(d, c, it); printf("MUFU.RSQ64H + DADD dependent: %.2f cyc/iter\n", rd());
  shfl_chain<<<1, 32>>>(d, c, it); printf("SHFL f64 dependent (2 SHFL): %.2f cyc/iter\n", rd());
  lds_chain<<<1, 32>>>(d, c, it); printf("LDS.64 + F2I dependent: %.2f cyc/iter\n", rd());
  seed_err<<<1, 32>>>(d); double he[64]; cudaMemcpy(he, d, sizeof(he), cudaMemcpyDeviceToHost);
  double wr = 0, wc = 0; for (int i = 0; i < 32; ++i) { wr = he[2*i] > wr ? he[2*i] : wr; wc = he[2*i+1] > wc ? he[2*i+1] : wc; }
  printf("seed rel. error: rsqrt.approx.ftz.f64 %.3e (2^%.1f), rcp.approx.ftz.f64 %.3e (2^%.1f)\n", wr, log2(wr), wc, log2(wc));
  return 0;
}
