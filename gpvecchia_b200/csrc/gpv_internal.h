// gpv_internal.h -- shared declarations of the library's translation units (not installed).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "u_kernels.cuh"

struct gpv_handle;
// Upper bound on the copy workers of one handle (results into pageable memory, gpv_capi.cu: pageable_copy); the
// multi-device front end divides the host's threads between its handles.
extern "C" void gpv_internal_set_copy_workers(gpv_handle* h, int n);

namespace gpv {

// One entry per compiled instantiation of u_sets_kernel<G,P,D>.
struct KernelEntry {
  int G, P, D;                                 // D = 0: runtime d <= GPV_MAX_D
  bool general;                                // general-nu table kernel vs closed forms
  int family;                                  // 0: two rows per lane (u_kernels.cuh), 1: band-folded, three or four rows per lane (u_band.cuh)
  const char* name;
  void (*kernel)(const UParams);
  int smem_bytes;
  int threads = 128;                           // block size (kThreadsPerBlock for families 0 and 1)
  int sets_per_block = 0;                      // sets per pass of the persistent loop; 0: kWarpsPerBlock * (32 / G)
};

// Picks the smallest instantiated P >= p for dimension d (the band-folded family wins a tie unless
// GPV_KERNEL_FAMILY=fold is set: development knob for A/B timing); nullptr if none.
const KernelEntry* select_kernel(int p, int d, bool general);

// Registration helpers implemented in u_inst_*.cu (one file per P so make -j parallelises nvcc).
void register_kernels_P4(KernelEntry* out, int* n);
void register_kernels_P8(KernelEntry* out, int* n);
void register_kernels_P11(KernelEntry* out, int* n);
void register_kernels_P16(KernelEntry* out, int* n);
void register_kernels_P21(KernelEntry* out, int* n);
void register_kernels_P26(KernelEntry* out, int* n);
void register_kernels_P31(KernelEntry* out, int* n);
void register_kernels_P32(KernelEntry* out, int* n);
void register_kernels_P41(KernelEntry* out, int* n);
void register_kernels_P51(KernelEntry* out, int* n);
void register_kernels_P64(KernelEntry* out, int* n);
void register_kernels_B8_21(KernelEntry* out, int* n);
void register_kernels_B8_26(KernelEntry* out, int* n);
void register_kernels_B8_31(KernelEntry* out, int* n);
void register_kernels_B8_32(KernelEntry* out, int* n);
void register_kernels_B16_41(KernelEntry* out, int* n);

}  // namespace gpv
