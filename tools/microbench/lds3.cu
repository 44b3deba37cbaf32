// Shared-memory wavefronts per warp instruction by access pattern; run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum
// (one launch per pattern, 1 warp, 1024 accesses): wavefronts / 1024 is the cost of the pattern.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 lds128(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, unsigned x) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(x), "r"(x)); }
__device__ unsigned pattern(int mode, int lane, int W) {
  switch (mode) {
    case 0: return 0;                                   // all lanes same chunk
    case 1: return (lane >> 4) * W;                     // 2 half-warps, adjacent chunks
    case 2: return (lane >> 4) * (2048 + 64);           // 2 half-warps, different lines, different banks
    case 3: return (lane >> 4) * 2048;                  // 2 half-warps, different lines, same banks
    case 4: return (lane >> 3) * W;                     // 4 quarter-warps, adjacent chunks
    case 5: return (lane >> 3) * (1024 + 32);           // 4 quarter-warps, different lines, different banks
    case 6: return (lane & 1) * W;                      // alternating lanes, 2 adjacent chunks
    case 7: return (lane & 15) * W;                     // 16 distinct adjacent chunks, both halves the same
    case 8: return lane * W;                            // every lane its own chunk
    case 9: return (lane >> 4) * 64;                    // 2 half-warps, same line, 64 B apart
    case 10: return (lane & 15) * 264;                  // 16 distinct, stride 33 doubles (conflict-free banks)
    case 11: return (lane & 15) * 256 + (lane >> 4) * 64; // 16 rows stride 32 doubles x 2 halves
    default: return 0;
  }
}
template <int W, bool ST>
__global__ void k(unsigned* out, int mode) {
  __shared__ __align__(128) unsigned sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + pattern(mode, lane, W);
  unsigned acc = 0;
  for (int i = 0; i < 128; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a = base + ((i * 8 + u) & 15) * 512;
      if (ST) sts64(a, acc + u);
      else if (W == 16) { uint4 v = lds128(a); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      else { uint2 v = lds64(a); acc ^= v.x ^ v.y; }
    }
  }
  out[threadIdx.x] = acc;
}
int main() {
  unsigned* d; cudaMalloc(&d, 1 << 20);
  for (int mode = 0; mode < 12; ++mode) { k<8, false><<<1, 32>>>(d, mode); cudaDeviceSynchronize(); }
  for (int mode = 0; mode < 12; ++mode) { k<16, false><<<1, 32>>>(d, mode); cudaDeviceSynchronize(); }
  for (int mode = 0; mode < 12; ++mode) { k<8, true><<<1, 32>>>(d, mode); cudaDeviceSynchronize(); }
  printf("done\n");
  return 0;
}
