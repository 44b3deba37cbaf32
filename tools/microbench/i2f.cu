// Does I2F.F64.S32 share the fp64 pipe with DFMA?  Kernel A: 8 independent DFMA chains per thread.  Kernel B: the
// same plus 2 int -> double conversions per 8 DFMAs (feeding a 9th, cheap chain).  Kernel C: 10 DFMAs per iteration
// (what B would cost if the conversions took DFMA slots).  nvcc -O3 -arch=sm_100a -o i2f i2f.cu && ./i2f
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double seed, int iseed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double e0 = a0 + 8, e1 = a0 + 9, s = 0.0;
  int n0 = iseed + threadIdx.x, n1 = iseed * 3 + threadIdx.x;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    if (MODE == 1) {
      s += 0.0;                        // keep the loop shape
      n0 = n0 * 1664525 + 1013904223; n1 = n1 * 22695477 + 1;
      double d0 = (double)n0, d1 = (double)n1;          // I2F.F64.S32 x 2
      asm volatile("" : "+d"(d0), "+d"(d1));
      e0 = __hiloint2double(__double2hiint(e0) ^ __double2hiint(d0), __double2loint(e1) ^ __double2loint(d1));
    } else if (MODE == 2) {
      e0 = fma(e0, b, c); e1 = fma(e1, b, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + e0 + e1 + s;
}
template <int MODE>
float run(double* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 256>>>(d, iters, 1.0, 7);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 256>>>(d, iters, 1.0, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* d; cudaMalloc(&d, sizeof(double) * 148 * 4 * 256);
  const int iters = 200000;
  float a = run<0>(d, iters), b = run<1>(d, iters), c = run<2>(d, iters);
  printf("8 DFMA: %.2f ms   8 DFMA + 2 I2F.F64: %.2f ms   10 DFMA: %.2f ms\n", a, b, c);
  printf("=> I2F.F64.S32 %s the fp64 pipe (extra cost per conversion = %.2f of a DFMA)\n",
         (b - a) < 0.35 * (c - a) ? "does NOT occupy" : "occupies", (b - a) / (c - a));
  return 0;
}
