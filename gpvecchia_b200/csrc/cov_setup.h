// cov_setup.h -- host-side set-up of the general-nu covariance table (CovTable, bessel_table.cuh): the constants
// of the direct evaluator and the range of w = d^2 the table covers.  Host code only; shared by gpv_capi.cu
// (setup_cov) and by the host emulation harness of the set kernels (tests/simt_emu).
#pragma once
#include <cmath>
#include "bessel_table.cuh"
#include "rgamma_coeffs.h"

namespace gpv {

inline void nu_constants(double nu, double sig2, CovTable* t) {
  // Temme's auxiliary functions for |mu| <= 1/2 from the Taylor series of 1/Gamma(1+x)
  const int nl = (int)(nu + 0.5);
  const double xmu = nu - nl;
  const double m2 = xmu * xmu;
  double g1 = 0.0, g2 = 0.0, pl = 0.0, mi = 0.0;
  // gam2 = sum_{k even} c_k mu^k ; gam1 = -sum_{k odd} c_k mu^(k-1)
  for (int k = GPV_RGAMMA_NCOEF - 1; k >= 0; --k) {
    pl = pl * xmu + kRGammaTaylor[k];
    mi = mi * (-xmu) + kRGammaTaylor[k];
  }
  for (int k = ((GPV_RGAMMA_NCOEF - 1) / 2) * 2; k >= 0; k -= 2) g2 = g2 * m2 + kRGammaTaylor[k];
  for (int k = ((GPV_RGAMMA_NCOEF - 2) / 2) * 2 + 1; k >= 1; k -= 2) g1 = g1 * m2 + kRGammaTaylor[k];
  t->nu = nu; t->nl = nl; t->xmu = xmu;
  t->gam1 = -g1; t->gam2 = g2; t->gampl = pl; t->gammi = mi;
  t->normcon = sig2 / (std::pow(2.0, nu - 1.0) * std::tgamma(nu));   // Matern.cpp:73
}

// table range: the top kTabOctaves octaves of w below the squared bounding-box diagonal w_max; for
// s = d / range >= kTabSSplit the table holds exp(+s) cov (w_split is an interval edge)
// win_top_exp: biased exponent of the highest octave of w the shared-memory window should hold (from the
// handle's histogram of neighbour distances), or < 0: the window ends at the top of the table.
constexpr int kTabWindow = GPV_TAB_WIN_OCT * kTabSub;   // intervals in the window (24 octaves of w: 4096 : 1 in distance)
inline void general_table_range(double range, double w_max, CovTable* tp, int win_top_exp = -1) {
  CovTable& t = *tp;
  double wmax = (w_max > 0.0 && std::isfinite(w_max)) ? w_max : 1.0;
  const int code_hi = hi32_of(wmax) >> (20 - kTabSubBits);
  int nint = kTabOctaves * kTabSub;
  int idx0 = code_hi - nint + 1;
  const int min_code = 1 << kTabSubBits;            // smallest normal exponent
  if (idx0 < min_code) { nint -= (min_code - idx0); idx0 = min_code; }
  t.idx0 = idx0; t.nint = nint; t.sub_bits = kTabSubBits; t.deg = kTabDeg;
  const double ws = (kTabSSplit * range) * (kTabSSplit * range);
  const int code_split = hi32_of(ws) >> (20 - kTabSubBits);
  t.w_split = from_hilo(code_split << (20 - kTabSubBits), 0);
  int win_end = idx0 + nint;                          // one past the last interval code of the window
  if (win_top_exp >= 0) {
    const int e = ((win_top_exp + 1) << kTabSubBits);
    if (e < win_end) win_end = e;
  }
  int win0 = win_end - kTabWindow - idx0;
  if (win0 > nint - kTabWindow) win0 = nint - kTabWindow;
  if (win0 < 0) win0 = 0;
  t.win0 = win0;
}

}  // namespace gpv
