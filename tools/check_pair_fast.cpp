// Host check of the pair stage's short sqrt / exp sequences (csrc/u_kernels.cuh: neg_sqrt_fast_n,
// exp_negarg_fast_n; compile-time switch GPV_PAIR_FAST).  The instruction sequences are restated with
// <cmath> fma (exact IEEE fused multiply-add, the same arithmetic DFMA performs); the one thing the host
// cannot reproduce is the MUFU.RSQ64H seed, so it is modelled by its measured envelope
// (tools/microbench/lat.cu: relative error <= 2^-19.9, low 32 bits zero) and swept over that envelope.
// Reference values: __float128 (libquadmath).  Prints the maximum errors in ulps of the RESULT for the
// sqrt and of 1 (absolute) / of the result (relative) for the exponential and for the Matern-1.5
// covariance as a function of the squared distance.
//   g++ -O2 -o /tmp/check_pair_fast tools/check_pair_fast.cpp -lquadmath && /tmp/check_pair_fast [samples]
// (tests/test_host_logic.py runs it with 200 000 samples per sweep.)
#include <quadmath.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>

static const double Q0 = 0.4999999999998501, Q1 = 0.16666666666664526, Q2 = 0.041666707476689775,
                    Q3 = 0.00833333916333604, K64LN2 = 92.33248261689366, LN2_64 = 0.010830424696249145,
                    LN2_64_HI = 0.010830424696223417, LN2_64_LO = 2.572804622327669e-14;
static double etab[64];

static uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static int hi(double x) { return (int)(bits(x) >> 32); }
static int lo(double x) { return (int)(bits(x) & 0xffffffffu); }
static double hilo(int h, int l) { return from_bits(((uint64_t)(uint32_t)h << 32) | (uint32_t)l); }

// seed model: 1/sqrt(w) (1 + delta), low word cleared
static double seed(double w, double delta) {
  const double y = (double)(1.0Q / sqrtq((__float128)w)) * (1.0 + delta);
  return from_bits(bits(y) & 0xffffffff00000000ull);
}
static double neg_sqrt_fast(double w, double delta) {
  const double y0 = seed(w, delta);
  const double g = w * (-y0);
  const double e = std::fma(g, y0, 1.0);
  const double a = std::fma(e, 0.375, 0.5), ge = g * e;
  return std::fma(a, ge, g);
}
static double sqrt_pos_old(double w, double delta) {   // the 9-instruction sequence it replaces (sqrt_pos_n)
  const double y0 = seed(w, delta);
  const double t = w * y0;
  const double e = std::fma(-t, y0, 1.0);
  const double y1 = std::fma(std::fma(e, 0.375, 0.5), e * y0, y0);
  const double g = w * y1, h = 0.5 * y1;
  const double r = std::fma(-g, h, 0.5);
  return std::fma(g, r, g);
}
static double exp_negarg(double ns_in, bool fast) {
  const double kShift = 6755399441055744.0;
  unsigned h = (unsigned)hi(ns_in);
  if (h > 0xC085E000u) h = 0xC085E000u;
  const double ns = hilo((int)h, lo(ns_in));
  const double t = std::fma(ns, K64LN2, kShift);
  const double kf = t - kShift;
  const int n = lo(t);
  double r;
  if (fast) r = std::fma(kf, -LN2_64, ns);
  else { r = std::fma(kf, -LN2_64_HI, ns); r = std::fma(kf, -LN2_64_LO, r); }
  const double T = etab[n & 63];
  const double r2 = r * r;
  double qq = std::fma(Q3, r, Q2);
  qq = std::fma(qq, r, Q1);
  qq = std::fma(qq, r, Q0);
  const double p = std::fma(qq, r2, r);
  const double v = std::fma(T, p, T);
  return hilo(hi(v) + (n >> 6) * 1048576, lo(v));
}
static double ulp_of(double x) { int e; std::frexp(x, &e); return std::ldexp(1.0, e - 53); }

int main(int argc, char** argv) {
  const int NIT = (argc > 1) ? std::atoi(argv[1]) : 2000000;
  for (int j = 0; j < 64; ++j) etab[j] = (double)powq(2.0Q, (__float128)j / 64);
  std::mt19937_64 rng(20240601);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const double dmax = std::ldexp(1.0, -20) * 1.08;   // 2^-19.9
  double worst_new = 0, worst_old = 0;
  for (int it = 0; it < NIT; ++it) {
    const double w = std::ldexp(1.0 + U(rng), (int)(U(rng) * 120) - 80);
    const double delta = (it % 3 == 0) ? dmax : (it % 3 == 1) ? -dmax : (2 * U(rng) - 1) * dmax;
    const __float128 ref = sqrtq((__float128)w);
    const double a = -neg_sqrt_fast(w, delta), b = sqrt_pos_old(w, delta);
    const double ua = (double)(fabsq((__float128)a - ref)) / ulp_of((double)ref);
    const double ub = (double)(fabsq((__float128)b - ref)) / ulp_of((double)ref);
    if (ua > worst_new) worst_new = ua;
    if (ub > worst_old) worst_old = ub;
  }
  std::printf("sqrt: max error %.3f ulp (fast, 5 instr)  %.3f ulp (old, 9 instr)\n", worst_new, worst_old);
  double wabs_f = 0, wabs_o = 0, wrel_f = 0, wrel_o = 0;
  for (int it = 0; it < NIT; ++it) {
    const double s = (it % 2) ? U(rng) * 40.0 : U(rng) * 700.0;
    const __float128 ref = expq(-(__float128)s);
    const double f = exp_negarg(-s, true), o = exp_negarg(-s, false);
    const double af = (double)fabsq((__float128)f - ref), ao = (double)fabsq((__float128)o - ref);
    wabs_f = std::fmax(wabs_f, af / std::ldexp(1.0, -53));
    wabs_o = std::fmax(wabs_o, ao / std::ldexp(1.0, -53));
    if (s < 40.0) {
      wrel_f = std::fmax(wrel_f, af / (double)ref / std::ldexp(1.0, -53));
      wrel_o = std::fmax(wrel_o, ao / (double)ref / std::ldexp(1.0, -53));
    }
  }
  std::printf("exp : max abs error %.3f (fast) %.3f (old) in units of 2^-53; max rel error for s < 40: %.2f (fast) %.2f (old) x 2^-53\n",
              wabs_f, wabs_o, wrel_f, wrel_o);
  // Matern 1.5 from the squared distance, c1 = sqrt(3)/range: error relative to c0 = 1
  double wcov = 0, wcov_old = 0;
  const double c1 = std::sqrt(3.0) / 0.004;
  for (int it = 0; it < NIT; ++it) {
    const double dist = std::ldexp(1.0 + U(rng), -(int)(U(rng) * 14) - 4);
    const double w = dist * dist + 1e-300;
    const double delta = (2 * U(rng) - 1) * dmax;
    const double ns = neg_sqrt_fast(w, delta) * c1;
    const double v = std::fma(-1.0, ns, 1.0) * exp_negarg(ns, true);
    const __float128 sq = sqrtq((__float128)w) * (__float128)c1;
    const __float128 ref = (1 + sq) * expq(-sq);
    wcov = std::fmax(wcov, (double)fabsq((__float128)v - ref) / std::ldexp(1.0, -53));
    const double nso = sqrt_pos_old(w, delta) * (-c1);
    const double vo = std::fma(-1.0, nso, 1.0) * exp_negarg(nso, false);
    wcov_old = std::fmax(wcov_old, (double)fabsq((__float128)vo - ref) / std::ldexp(1.0, -53));
  }
  std::printf("matern 1.5 covariance (c0 = 1): max abs error %.3f (fast) %.3f (old) x 2^-53\n", wcov, wcov_old);
  // special arguments: guard value, clamp, NaN
  const double z = neg_sqrt_fast(1e-300, 0.0);
  std::printf("special: -sqrt(1e-300) = %.3e, exp(ns = -1e-150) = %.17g, exp(-1e9) = %.3e, exp(NaN) = %f\n", z,
              exp_negarg(z, true), exp_negarg(-1e9, true), exp_negarg(std::nan(""), true));
  const bool ok = worst_new < 1.0 && wabs_f < 1.25 && wcov < 1.5 * wcov_old && exp_negarg(z, true) == 1.0 &&
                  std::isnan(exp_negarg(std::nan(""), true)) && exp_negarg(-1e9, true) < 1e-300;
  std::printf(ok ? "OK\n" : "FAIL\n");
  return ok ? 0 : 1;
}
