// Shared-memory load throughput by access pattern (integer accumulate, so the LSU is the limiter).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 lds128(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
// mode: 0 all lanes same 16B; 1 two half-warps, adjacent 16B chunks (same 128B line); 2 two half-warps, different
// lines/banks (+2064 B); 3 four quarter-warps adjacent chunks; 4 each lane own 16B (512 B contiguous)
template <int W>
__global__ void k(unsigned* out, long long* cyc, int iters, int mode) {
  __shared__ __align__(128) unsigned sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned off = 0;
  if (mode == 1) off = (lane >> 4) * W;
  if (mode == 2) off = (lane >> 4) * (2048 + 64);
  if (mode == 3) off = (lane >> 3) * W;
  if (mode == 4) off = lane * W;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + off + (threadIdx.x >> 5) * 1024;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a = base + ((i * 8 + u) & 15) * 32;
      if (W == 16) { uint4 v = lds128(a); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      else { uint2 v = lds64(a); acc ^= v.x ^ v.y; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  unsigned* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
  long long h; const int it = 2048;
  const char* names[] = {"all lanes same chunk", "2 halves, adjacent chunks (same line)", "2 halves, different lines", "4 quarters adjacent chunks", "every lane its own chunk"};
  for (int W : {8, 16}) for (int mode = 0; mode < 5; ++mode) {
    if (W == 16) k<16><<<1, 512>>>(d, c, it, mode); else k<8><<<1, 512>>>(d, c, it, mode);
    cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("LDS.%-3d %-40s: %.2f SM-cycles per warp instruction\n", W * 8, names[mode], (double)h / (it * 8.0) / 16.0);
  }
  return 0;
}
