// gpv_multi.cu -- one-process, many-GPU front end of the C ABI (include/gpvecchia_b200.h).
//
// An R session is a single process, so the reference-facing way to use the 8 GPUs of a box is one
// handle per device driven from worker threads: rows are independent (src/U_NZentries.cpp:39-69; the
// reference uses schedule(static)), each device owns a contiguous row range balanced by sum n0^3,
// writes its slice of the packed createU.R:158-160 vector straight into the caller's buffer, and the
// four likelihood partial sums are added on the host.  No inter-GPU data path exists, hence no NCCL
// here; the multi-process form (bench.py under torchrun) combines the same partial sums with one
// all-reduce.
#include "../../include/gpvecchia_b200.h"
extern "C" void gpv_internal_set_copy_workers(gpv_handle* h, int n);   // gpv_capi.cu

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct gpv_multi {
  std::vector<gpv_handle*> h;
  std::vector<int> dev;
  std::vector<int64_t> cut;        // row cuts, size ndev + 1
  std::vector<int64_t> off;        // packed offsets, size ndev + 1
  int64_t Nlocs = 0, n_obs = 0;
  int p = 0;
  // NVLink exchange between the devices of this process (gpv_dist_*): each device uploads 1/ndev of the per-call
  // vectors of a likelihood call and NCCL broadcasts fill the rest, instead of ndev full uploads through one host
  bool dist = false;
  std::vector<int64_t> loc_cuts, obs_cuts;
};

extern "C" void gpv_set_last_error(const char* msg);   // gpv_capi.cu

namespace {
struct Result { gpv_status st = GPV_OK; std::string msg; };

template <class F>
gpv_status run_all(int n, F f) {
  std::vector<Result> res(n);
  std::vector<std::thread> th;
  th.reserve(n);
  for (int i = 0; i < n; ++i)
    th.emplace_back([&, i]() {
      res[i].st = f(i);
      if (res[i].st != GPV_OK) res[i].msg = gpv_last_error();   // thread-local in the worker
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (res[i].st != GPV_OK) {
      gpv_set_last_error(("device shard " + std::to_string(i) + ": " + res[i].msg).c_str());
      return res[i].st;
    }
  return GPV_OK;
}
}  // namespace

extern "C" gpv_status gpv_multi_create(gpv_multi** out, int64_t Nlocs, int p, int d, const double* locs,
                                       const int32_t* revNNarray, const void* revCond,
                                       gpv_cond_type cond_type, const int32_t* obs, const int* devices,
                                       int ndev) {
  if (!out) { gpv_set_last_error("gpv_multi_create: out is null"); return GPV_ERR_ARG; }
  *out = nullptr;
  if (!devices || ndev < 1 || !revNNarray || Nlocs <= 0 || p <= 0) {
    gpv_set_last_error("gpv_multi_create: bad argument");
    return GPV_ERR_ARG;
  }
  gpv_multi* m = new gpv_multi();
  m->Nlocs = Nlocs; m->p = p;
  m->dev.assign(devices, devices + ndev);
  // contiguous cuts balancing sum n0^3 (the factorisation cost; n0 = 1 rows of `zy` layouts are free)
  std::vector<double> w((size_t)Nlocs + 1, 0.0);
  for (int64_t r = 0; r < Nlocs; ++r) {
    int n0 = 0;
    for (int j = 0; j < p; ++j) n0 += revNNarray[r + (int64_t)j * Nlocs] > 0;   // 0, NA (INT_MIN): missing
    w[r + 1] = w[r] + (double)n0 * n0 * n0;
  }
  m->cut.assign(ndev + 1, 0);
  m->cut[ndev] = Nlocs;
  int64_t r = 0;
  for (int k = 1; k < ndev; ++k) {
    const double target = w[Nlocs] * k / ndev;
    while (r < Nlocs && w[r + 1] < target) ++r;
    m->cut[k] = r < m->cut[k - 1] ? m->cut[k - 1] : r;
  }
  m->h.assign(ndev, nullptr);
  gpv_status st = run_all(ndev, [&](int i) {
    return gpv_create(&m->h[i], Nlocs, p, d, locs, revNNarray, revCond, cond_type, obs, m->cut[i],
                      m->cut[i + 1], m->dev[i]);
  });
  if (st != GPV_OK) {
    for (auto* hh : m->h) gpv_destroy(hh);
    delete m;
    return st;
  }
  m->off.assign(ndev + 1, 0);
  for (int i = 0; i < ndev; ++i) m->off[i + 1] = m->off[i] + gpv_packed_len(m->h[i]);
  {
    // results into pageable memory: the handles' copy workers share the host's threads
    const unsigned hc = std::thread::hardware_concurrency();
    const int per = (int)((hc ? hc : 8u) / (2u * (unsigned)ndev));
    for (int i = 0; i < ndev; ++i) gpv_internal_set_copy_workers(m->h[i], per < 2 ? 2 : per);
  }
  if (obs) for (int64_t i = 0; i < Nlocs; ++i) m->n_obs += (obs[i] != 0 && obs[i] != INT32_MIN);
  // one NCCL communicator over the devices (distinct ordinals, no empty shard): optional, the host path stays
  bool distinct = ndev > 1 && m->n_obs > 0;
  for (int i = 0; i < ndev && distinct; ++i) {
    if (m->cut[i + 1] <= m->cut[i]) distinct = false;
    for (int j = 0; j < i; ++j) if (m->dev[j] == m->dev[i]) distinct = false;
  }
  if (distinct) {
    char uid[128];
    if (gpv_dist_unique_id(uid) == GPV_OK &&
        run_all(ndev, [&](int i) { return gpv_dist_init(m->h[i], uid, i, ndev); }) == GPV_OK) {
      m->dist = true;
      m->loc_cuts.resize(ndev + 1);
      m->obs_cuts.resize(ndev + 1);
      for (int i = 0; i <= ndev; ++i) { m->loc_cuts[i] = Nlocs * i / ndev; m->obs_cuts[i] = m->n_obs * i / ndev; }
    } else {
      for (auto* hh : m->h) gpv_dist_finalize(hh);
    }
  }
  *out = m;
  return GPV_OK;
}

extern "C" void gpv_multi_destroy(gpv_multi* m) {
  if (!m) return;
  for (auto* hh : m->h) gpv_destroy(hh);
  delete m;
}
extern "C" int gpv_multi_num_devices(const gpv_multi* m) { return m ? (int)m->h.size() : 0; }
extern "C" int64_t gpv_multi_packed_len(const gpv_multi* m) { return m ? m->off.back() : 0; }
extern "C" void gpv_multi_row_cuts(const gpv_multi* m, int64_t* cuts) {
  if (m && cuts) std::memcpy(cuts, m->cut.data(), sizeof(int64_t) * m->cut.size());
}

extern "C" gpv_status gpv_multi_set_revcond(gpv_multi* m, const void* revCond, gpv_cond_type cond_type) {
  if (!m) { gpv_set_last_error("gpv_multi_set_revcond: null handle"); return GPV_ERR_ARG; }
  return run_all((int)m->h.size(), [&](int i) { return gpv_set_revcond(m->h[i], revCond, cond_type); });
}

extern "C" gpv_status gpv_multi_u_values_packed(gpv_multi* m, const char* covType, const double* covparms,
                                                int ncov, const double* nuggets,
                                                const double* nuggets_obsord, int64_t n,
                                                int zentries_tail, double* out, int64_t* nfail,
                                                int64_t* first_fail) {
  if (!m || !out) { gpv_set_last_error("gpv_multi_u_values_packed: null argument"); return GPV_ERR_ARG; }
  const int nd = (int)m->h.size();
  std::vector<int64_t> nf(nd, 0), ff(nd, -1);
  // shard i writes its rows' values at out + off[i]; the last shard also appends the Z values, which
  // land right behind its own (= the global) packed block
  gpv_status st = run_all(nd, [&](int i) {
    const int tail = (zentries_tail && i == nd - 1) ? 1 : 0;
    return gpv_u_values_packed(m->h[i], covType, covparms, ncov, nuggets, nuggets_obsord, n, tail,
                               out + m->off[i], &nf[i], &ff[i]);
  });
  if (st != GPV_OK) return st;
  int64_t tot = 0, first = -1;
  for (int i = 0; i < nd; ++i) {
    tot += nf[i];
    if (nf[i] > 0 && (first < 0 || ff[i] < first)) first = ff[i];
  }
  if (nfail) *nfail = tot;
  if (first_fail) *first_fail = first;
  return GPV_OK;
}

// ---- compressed-column output over several devices -------------------------------------------------
// The U columns of consecutive row shards are consecutive column ranges, so every device writes its
// slice of dgCMatrix@i / @x in place; the column pointers of shard i are shifted by the nonzeros before.
static gpv_status multi_csc_offsets(gpv_multi* m, std::vector<int64_t>* col0, std::vector<int64_t>* nz0, int64_t* size) {
  const int nd = (int)m->h.size();
  col0->assign(nd + 1, 0); nz0->assign(nd + 1, 0);
  for (int i = 0; i < nd; ++i) {
    int64_t nc = 0, nz = 0, sz = 0;
    gpv_status st = gpv_csc_dims(m->h[i], &nc, &nz, &sz);
    if (st != GPV_OK) return st;
    (*col0)[i + 1] = (*col0)[i] + nc;
    (*nz0)[i + 1] = (*nz0)[i] + nz;
    if (size) *size = sz;
  }
  return GPV_OK;
}

extern "C" gpv_status gpv_multi_csc_dims(gpv_multi* m, int64_t* ncols, int64_t* nnz, int64_t* size) {
  if (!m) { gpv_set_last_error("gpv_multi_csc_dims: null handle"); return GPV_ERR_ARG; }
  std::vector<int64_t> c0, z0;
  gpv_status st = multi_csc_offsets(m, &c0, &z0, size);
  if (st != GPV_OK) return st;
  if (ncols) *ncols = c0.back();
  if (nnz) *nnz = z0.back();
  return GPV_OK;
}

extern "C" gpv_status gpv_multi_u_csc_pattern(gpv_multi* m, int32_t* colptr, int32_t* rowidx) {
  if (!m || !colptr || !rowidx) { gpv_set_last_error("gpv_multi_u_csc_pattern: null argument"); return GPV_ERR_ARG; }
  std::vector<int64_t> c0, z0;
  gpv_status st = multi_csc_offsets(m, &c0, &z0, nullptr);
  if (st != GPV_OK) return st;
  if (z0.back() > 2147483647LL) { gpv_set_last_error("matrix too large for 32-bit dgCMatrix indices"); return GPV_ERR_UNSUPPORTED; }
  const int nd = (int)m->h.size();
  // shard i fills colptr[c0[i] .. c0[i+1]] relative to its own first nonzero; the shared boundary entry
  // is rewritten by the next shard with the same value after the shift, so shift from the last shard down
  st = run_all(nd, [&](int i) {
    std::vector<int32_t> cp((size_t)(c0[i + 1] - c0[i] + 1));
    gpv_status s2 = gpv_u_csc_pattern(m->h[i], cp.data(), rowidx + z0[i]);
    if (s2 != GPV_OK) return s2;
    for (int64_t c = 0; c < c0[i + 1] - c0[i]; ++c) colptr[c0[i] + c] = (int32_t)(cp[(size_t)c] + z0[i]);
    if (i == nd - 1) colptr[c0[nd]] = (int32_t)z0[nd];
    return GPV_OK;
  });
  return st;
}

extern "C" gpv_status gpv_multi_u_values_csc(gpv_multi* m, const char* covType, const double* covparms, int ncov,
                                             const double* nuggets, const double* nuggets_obsord, int64_t n,
                                             double* x, int64_t* nfail, int64_t* first_fail) {
  if (!m || !x) { gpv_set_last_error("gpv_multi_u_values_csc: null argument"); return GPV_ERR_ARG; }
  std::vector<int64_t> c0, z0;
  gpv_status st = multi_csc_offsets(m, &c0, &z0, nullptr);
  if (st != GPV_OK) return st;
  const int nd = (int)m->h.size();
  std::vector<int64_t> nf(nd, 0), ff(nd, -1);
  st = run_all(nd, [&](int i) {
    return gpv_u_values_csc(m->h[i], covType, covparms, ncov, nuggets, nuggets_obsord, n, x + z0[i], &nf[i], &ff[i]);
  });
  if (st != GPV_OK) return st;
  int64_t tot = 0, first = -1;
  for (int i = 0; i < nd; ++i) {
    tot += nf[i];
    if (nf[i] > 0 && (first < 0 || ff[i] < first)) first = ff[i];
  }
  if (nfail) *nfail = tot;
  if (first_fail) *first_fail = first;
  return GPV_OK;
}

extern "C" gpv_status gpv_multi_loglik_numerator(gpv_multi* m, const char* covType, const double* covparms,
                                                 int ncov, const double* nuggets,
                                                 const double* nuggets_obsord, const double* zord,
                                                 int64_t n, int64_t skip_rows, double out[3]) {
  if (!m || !out) { gpv_set_last_error("gpv_multi_loglik_numerator: null argument"); return GPV_ERR_ARG; }
  const int nd = (int)m->h.size();
  std::vector<double> parts(3 * (size_t)nd, 0.0);
  gpv_status st = run_all(nd, [&](int i) {
    return gpv_loglik_numerator(m->h[i], covType, covparms, ncov, nuggets, nuggets_obsord, zord, n,
                                skip_rows, i == 0 ? 1 : 0, &parts[3 * (size_t)i]);
  });
  if (st != GPV_OK) return st;
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < nd; ++i)            // fixed order: reproducible
    for (int k = 0; k < 3; ++k) out[k] += parts[3 * (size_t)i + k];
  return GPV_OK;
}

extern "C" gpv_status gpv_multi_loglik_z(gpv_multi* m, const char* covType, const double* covparms, int ncov,
                                         const double* nuggets, const double* nuggets_obsord,
                                         const double* zord, int64_t n, double out[6]) {
  if (!m || !out) { gpv_set_last_error("gpv_multi_loglik_z: null argument"); return GPV_ERR_ARG; }
  const int nd = (int)m->h.size();
  std::vector<double> parts(6 * (size_t)nd, 0.0);
  if (m->dist && nuggets && nuggets_obsord && zord && n == m->n_obs) {
    // every device uploads its slice, the slices travel over NVLink, the partial sums are all-reduced on the
    // devices: each shard returns the complete result
    gpv_status sd = run_all(nd, [&](int i) {
      return gpv_loglik_z_dist(m->h[i], covType, covparms, ncov, nuggets + m->loc_cuts[i], nuggets_obsord + m->obs_cuts[i],
                               zord + m->obs_cuts[i], m->loc_cuts.data(), m->obs_cuts.data(), &parts[6 * (size_t)i]);
    });
    if (sd != GPV_OK) return sd;
    for (int k = 0; k < 6; ++k) out[k] = parts[k];
    return GPV_OK;
  }
  gpv_status st = run_all(nd, [&](int i) {
    return gpv_loglik_z(m->h[i], covType, covparms, ncov, nuggets, nuggets_obsord, zord, n, i == 0 ? 1 : 0,
                        &parts[6 * (size_t)i]);
  });
  if (st != GPV_OK) return st;
  for (int k = 1; k < 6; ++k) {
    out[k] = 0.0;
    for (int i = 0; i < nd; ++i) out[k] += parts[6 * (size_t)i + k];
  }
  const double neg2 = out[2] - out[4] + out[1] - out[3] + (double)n * std::log(2.0 * 3.141592653589793238462643383279502884);
  out[0] = -0.5 * neg2;
  return GPV_OK;
}
