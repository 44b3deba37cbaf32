// tests/simt_emu/tsan_main.cpp -- TEST INFRASTRUCTURE: drives emu_harness.cpp under ThreadSanitizer.  With one
// host thread per CUDA thread and barriers only where the kernel synchronises, a shared-memory exchange that is
// not ordered by a __syncwarp / __syncthreads is a data race that TSAN reports (racecheck on the host).
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -Itests/simt_emu -Igpvecchia_b200/csrc -Iinclude \
//       tests/simt_emu/emu_harness.cpp tests/simt_emu/tsan_main.cpp -o /tmp/emu_tsan && /tmp/emu_tsan
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

extern "C" int emu_u_sets(int family, int G, int P, int D, int grid, int64_t nsets, int p, int d, const double* locs,
                          const int32_t* nn, const uint64_t* cond, const double* nuggets, double* out,
                          const int64_t* row_off, const double* zloc, int full_z, double* partials,
                          unsigned long long* nfail, long long* first_fail, int cov, const double* c);

static int run(int family, int G, int P, int d, int n) {
  const int p = P;
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<double> locs((size_t)n * d), nug(n, 0.1), z(n), out((size_t)n * p, NAN), part(2 * 4, NAN);
  for (auto& v : locs) v = U(rng);
  for (auto& v : z) v = U(rng) - 0.5;
  std::vector<int32_t> nn((size_t)n * p, -1);
  std::vector<uint64_t> cond(n, 0);
  for (int k = 0; k < n; ++k) {            // the p - 1 previous points (any earlier points make a valid set), self last
    const int n0 = std::min(k + 1, p);
    for (int j = 0; j < n0; ++j) nn[(size_t)k * p + (p - n0) + j] = k - (n0 - 1) + j;
    cond[k] = 1ull << (p - 1);
  }
  unsigned long long nfail = 0;
  long long first = INT64_MAX;
  const double c[5] = {1.0, std::sqrt(3.0) / 0.3, 0, 0, 0};
  const int rc = emu_u_sets(family, G, P, d, 2, n, p, d, locs.data(), nn.data(), cond.data(), nug.data(), out.data(), nullptr,
                            z.data(), 1, part.data(), &nfail, &first, 1, c);
  double s = 0;
  for (double v : out) s += v;
  std::printf("family=%d G=%d P=%d d=%d: rc=%d nfail=%llu checksum=%.12g partial0=%.12g\n", family, G, P, d, rc, nfail, s, part[0] + part[4]);
  return rc != 0 || nfail != 0 || !(s == s);
}

int main() {
  int bad = run(1, 8, 31, 2, 70);
  bad |= run(1, 16, 41, 3, 40);
  bad |= run(0, 16, 31, 2, 40);    // two-row kernel, two sets per warp
  bad |= run(0, 32, 51, 2, 20);    // two-row kernel, one set per warp
  bad |= run(1, 8, 21, 3, 60);     // three bands of eight lanes
  return bad;
}
