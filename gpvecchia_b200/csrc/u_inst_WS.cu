// u_inst_WS.cu -- the warp-specialised experiment (u_band_ws.cuh): G = 8, d = 2; closed forms and, for P = 31, general nu.  Registered as
// kernel family 2, which select_kernel only returns when GPV_KERNEL_FAMILY=ws is set (development knob).
#include "gpv_internal.h"
#include "bessel_table.cuh"
#include "u_band_ws.cuh"

namespace gpv {

template <int P, int D, bool GENERAL>
static KernelEntry make_ws_entry(const char* name) {
  KernelEntry e;
  e.G = 8; e.P = P; e.D = D; e.general = GENERAL; e.family = 2; e.name = name;
  e.kernel = u_band_ws_kernel<P, D, GENERAL>;
  e.smem_bytes = WsLayout<P, D>::kBytesPerBlock;
  e.threads = kWsThreads;
  e.sets_per_block = WsLayout<P, D>::kSetsPerBlock;
  return e;
}

void register_kernels_WS(KernelEntry* out, int* n) {
  out[(*n)++] = make_ws_entry<26, 2, false>("u_band_ws<P=26,D=2,closed>");
  out[(*n)++] = make_ws_entry<31, 2, false>("u_band_ws<P=31,D=2,closed>");
  out[(*n)++] = make_ws_entry<32, 2, false>("u_band_ws<P=32,D=2,closed>");
  out[(*n)++] = make_ws_entry<31, 2, true>("u_band_ws<P=31,D=2,general>");
}

}  // namespace gpv
