// host_specify.cu -- host-side input producer that belongs to the path's callers:
// whichCondOnLatent (R/whichCondOnLatent.R:2-27), the sparse-general-Vecchia choice of which
// neighbours to condition on as latent y.  Row k depends on the finished rows of its neighbours, so it
// is a sequential sweep like the reference's R loop; what changes is the cost per row: sorted id lists
// and merge counts, O(m^2) instead of R's O(m^3) `is.element` calls through the interpreter
// (n = 1e6, m = 30: about a second instead of hours).  No device work: this is not the hot path and
// has no CUDA variant.
#include "../../include/gpvecchia_b200.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <vector>

extern "C" void gpv_set_last_error(const char* msg);   // gpv_capi.cu

namespace {
inline bool is_na(int32_t v) { return v == INT_MIN || v <= 0; }   // NA_integer_, or the 0 of revNNarray
// |{a in A} ∩ {b in B}| for two ascending lists without duplicates
inline int merge_count(const int32_t* a, int na, const int32_t* b, int nb) {
  int i = 0, j = 0, c = 0;
  while (i < na && j < nb) {
    if (a[i] < b[j]) ++i;
    else if (a[i] > b[j]) ++j;
    else { ++c; ++i; ++j; }
  }
  return c;
}
}  // namespace

extern "C" gpv_status gpv_whichCondOnLatent(const int32_t* NNarray, int64_t n, int p, int64_t firstind_pred,
                                            int32_t* CondOnLatent) {
  if (!NNarray || !CondOnLatent || n <= 0 || p <= 0) { gpv_set_last_error("gpv_whichCondOnLatent: bad argument"); return GPV_ERR_ARG; }
  if (firstind_pred <= 0) firstind_pred = n + 1;                      // default of the R signature
  const size_t N = (size_t)n;
  for (size_t i = 0; i < N * (size_t)p; ++i) CondOnLatent[i] = INT_MIN;   // matrix(NA, n, m+1)
  // ascending list of the ids a finished row conditions on as latent: NNarray[l,] * CondOnLatent[l,]
  // without its NA and 0 entries (an id is never 0, so the FALSE entries cannot match anything)
  std::vector<int32_t> tids(N * (size_t)p);
  std::vector<int32_t> tcnt(N, 0), nna(N, 0);
  std::vector<int32_t> row(p), srt(p), lat(p);
  auto at = [&](size_t k, int j) { return NNarray[k + N * (size_t)j]; };
  auto finish = [&](size_t k) {
    int c = 0, na = 0;
    for (int j = 0; j < p; ++j) {
      const int32_t id = at(k, j);
      if (is_na(id)) { ++na; continue; }
      if (CondOnLatent[k + N * (size_t)j] == 1) tids[k * (size_t)p + c++] = id;
    }
    std::sort(tids.begin() + k * (size_t)p, tids.begin() + k * (size_t)p + c);
    c = (int)(std::unique(tids.begin() + k * (size_t)p, tids.begin() + k * (size_t)p + c) - (tids.begin() + k * (size_t)p));
    tcnt[k] = c;
    nna[k] = na;                                                       // NA ids (their products are NA too)
  };
  CondOnLatent[0] = 1;                                                 // CondOnLatent[1,1] <- TRUE
  finish(0);
  for (size_t k = 1; k < N; ++k) {
    int ns = 0, na_k = 0;
    for (int j = 0; j < p; ++j) {
      row[j] = at(k, j);
      if (is_na(row[j])) ++na_k; else srt[ns++] = row[j];
    }
    std::sort(srt.begin(), srt.begin() + ns);
    // is.element counts every element of NNarray[k,] that is in the table: duplicates in the row count
    // twice (they do not occur in neighbour arrays; kept exact by counting on the unsorted row below)
    const bool dup = std::adjacent_find(srt.begin(), srt.begin() + ns) != srt.begin() + ns;
    lat[0] = 0;                                                        // latents = rep(0, m), extended to m+1
    for (int ind = 1; ind < p; ++ind) {
      lat[ind] = 0;
      const int32_t l = row[ind];
      if (is_na(l) || (int64_t)l >= firstind_pred || (int64_t)l > n) continue;
      const size_t lr = (size_t)l - 1;
      const int32_t* T = tids.data() + lr * (size_t)p;
      int c;
      if (!dup) {
        c = merge_count(srt.data(), ns, T, tcnt[lr]);
      } else {
        c = 0;
        for (int j = 0; j < p; ++j)
          if (!is_na(row[j]) && std::binary_search(T, T + tcnt[lr], row[j])) ++c;
      }
      // match() pairs an NA of the row with an NA of the table: every NA of row k counts once the table has one
      if (na_k > 0 && (nna[lr] > 0 || lr >= k)) c += na_k;
      lat[ind] = c;
    }
    int best = 0;
    for (int ind = 1; ind < p; ++ind) if (lat[ind] > lat[best]) best = ind;   // which(latents == max)[1]
    const int32_t chosen = row[best];
    const int32_t* T = nullptr;
    int nt = 0;
    if (!is_na(chosen) && (int64_t)chosen <= n && (size_t)chosen - 1 < k) { T = tids.data() + ((size_t)chosen - 1) * (size_t)p; nt = tcnt[(size_t)chosen - 1]; }
    for (int j = 0; j < p; ++j) {
      if (is_na(row[j])) continue;                                     // stays NA (:22)
      int v = (nt > 0 && std::binary_search(T, T + nt, row[j])) ? 1 : 0;
      if ((int64_t)row[j] >= firstind_pred) v = 1;                     // :20
      CondOnLatent[k + N * (size_t)j] = v;
    }
    if (!is_na(row[0])) CondOnLatent[k] = 1;                           // :21
    finish(k);
  }
  return GPV_OK;
}
