/* gpvecchia_b200.h -- C ABI of the B200-native U_NZentries / likelihood-numerator path.
 *
 * Drop-in boundary: these entry points are what the reference's `.Call` glue for this path
 * binds.  Reference interface replaced (GPvecchia 0.1.8):
 *   R stub      U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets,
 *               nuggets_obsord, covType, covparms)             R/RcppExports.R:22-24
 *   C++ glue    _GPvecchia_U_NZentries(9 SEXP)                 src/RcppExports.cpp:49-67
 *   C++ kernel  U_NZentries(...)                               src/U_NZentries.cpp:25-118
 *   consumers   createU()                                      R/createU.R:141-163
 *               vecchia_likelihood_U() numerator               R/vecchia_likelihood.R:71-76
 *
 * Conventions
 *   - plain C, no exceptions, no torch/R types; every function returns a gpv_status;
 *     gpv_last_error() gives the message of the last failure on the calling thread.
 *   - all host matrices are COLUMN-MAJOR (R layout): locs[k + N*c], revNN[k + N*j].
 *   - revNNarray: int32, 1-based location ids, 0 = missing (createU.R:146-147); missing
 *     entries are compacted exactly as `inds.elem(find(inds))` does (U_NZentries.cpp:44).
 *   - revCondOnLatent: either R's logical storage (int32: 1 TRUE, 0 FALSE, INT_MIN NA) or the
 *     coerced double the reference kernel sees (1.0 / 0.0 / NaN): selected by gpv_cond_type.
 *   - the caller owns every host buffer before and after each call; the library never keeps a
 *     host pointer past return.  Device memory, streams and events belong to the handle.
 *   - one thread at a time per handle.  Calls block until outputs are in host memory
 *     (the *_dev variants are asynchronous on the handle's stream unless stated).
 *   - a non positive-definite block is NOT an error (U_NZentries.cpp:60-66): its row is left
 *     zero and counted in *nfail; *first_fail is the smallest failing 0-based row or -1.
 */
#ifndef GPVECCHIA_B200_H
#define GPVECCHIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpv_handle gpv_handle;

typedef enum {
  GPV_OK = 0,
  GPV_ERR_ARG = 1,         /* bad argument (null pointer, size, p > GPV_MAX_P, ...) */
  GPV_ERR_CUDA = 2,        /* CUDA runtime failure; message has the CUDA error string */
  GPV_ERR_COVTYPE = 3,     /* covType not "matern"/"esqe" (the reference only prints, :27-29) */
  GPV_ERR_NOMEM = 4,
  GPV_ERR_UNSUPPORTED = 5  /* valid input outside what this build instantiates */
} gpv_status;

typedef enum { GPV_COND_RLOGICAL_I32 = 0, GPV_COND_F64 = 1 } gpv_cond_type;

/* covType strings of the reference: "matern" -> covparms (sig2, range, nu), "esqe" ->
 * (sig2_1, r1, sig2_2, r2).  Matern picks the closed form by exact == on nu (Matern.cpp:32,43,58)
 * and otherwise the general K_nu branch (:72-83, no sqrt(2 nu) scaling). */
#define GPV_MAX_P 64
#define GPV_MAX_D 8

const char* gpv_last_error(void);
const char* gpv_version(void);
int gpv_device_count(void);
/* Page-locked (pinned, portable) host memory: a result buffer allocated here gets the overlapped copy pipeline of the
 * U-values calls (5 ms instead of 13 ms for the 264 MB of n = 1e6, m = 30).  NULL if it cannot be had.  Used by the
 * R shim's optional custom allocator (Rf_allocVector3), INTEGRATION.md. */
void* gpv_host_alloc(size_t bytes);
void gpv_host_free(void* p);

/* ---- handle: uploads the parameter-free arrays of one vecchia.approx once ------------------
 * Nlocs, p=m+1, d : shapes of locsord (Nlocs x d) and revNNarray/revCond (Nlocs x p).
 * obs            : Nlocs R-logical int32 (vecchia.approx$obs, in locsord order), may be NULL
 *                  when the likelihood numerator is not needed.
 * row_begin/end  : this handle computes rows [row_begin,row_end) only (multi-GPU sharding by
 *                  contiguous row range; pass 0,Nlocs for everything).  locs and obs are always
 *                  uploaded whole (neighbour ids of late rows span the full index range).
 * device         : CUDA device ordinal. */
gpv_status gpv_create(gpv_handle** out, int64_t Nlocs, int p, int d, const double* locs,
                      const int32_t* revNNarray, const void* revCondOnLatent,
                      gpv_cond_type cond_type, const int32_t* obs, int64_t row_begin,
                      int64_t row_end, int device);
/* Same, but revNN_rows / revCond_rows hold ONLY the rows [row_begin,row_end) (column-major with
 * leading dimension row_end - row_begin): what a rank of a multi-GPU job has after sharding. */
gpv_status gpv_create_shard(gpv_handle** out, int64_t Nlocs, int p, int d, const double* locs,
                            const int32_t* revNN_rows, const void* revCond_rows,
                            gpv_cond_type cond_type, const int32_t* obs, int64_t row_begin,
                            int64_t row_end, int device);
void gpv_destroy(gpv_handle* h);

/* Re-upload revCond (createU.R:83-86 rewrites it per call when some nuggets are zero). */
gpv_status gpv_set_revcond(gpv_handle* h, const void* revCondOnLatent, gpv_cond_type cond_type);

/* ---- U_NZentries through a handle (host buffers) --------------------------------------------
 * nuggets[Nlocs] (nuggets.all.ord), nuggets_obsord[n].  Lentries: column-major (row_end-
 * row_begin) x p, zero-filled beyond n0 like the reference; Zentries[2n] or NULL.
 * A whole-range handle takes the nuggets of all its observations (n = sum(obs)).  A row shard may be given
 * the nuggets of any contiguous slice of the observations (n <= sum(obs)) and returns THEIR Zentries
 * (elementwise, U_NZentries.cpp:110-115): the ranks of a sharded run each move only their slice.
 * revNNarray ids: 1-based; 0, NA_integer_ and any non-positive value are "missing"; an id > Nlocs makes
 * gpv_create fail with GPV_ERR_ARG.
 * Host output buffers: page-locked memory gets the overlapped (chunked) copy pipeline on a copy stream.  Pageable
 * memory (every R vector) of 8 MB or more gets the same chunked launches, and up to 8 worker threads of the
 * library fetch 4 MB pieces into page-locked slots and copy them into the caller's buffer, so that the page
 * faults of a freshly allocated vector and the host copies run in parallel (n = 1e6, m = 30: 13 ms instead of
 * the 60 ms of one cudaMemcpy into fresh pages; 5 ms page-locked).  The workers touch the two buffers only;
 * GPV_COPY_WORKERS=1..16 overrides their number (default: half the host's hardware threads, at most 8).
 * Smaller pageable outputs: one launch and one copy.
 * Of `nuggets` a call reads entries [0, gpv_nuggets_read(h)): one past the largest id the handle's rows name
 * (all Nlocs for a whole-range handle; a prefix for a row shard of an ordered layout), and that is all it
 * uploads.  In the chunked pipeline the upload is staged with the chunks (chunk c brings up what its rows name
 * and no earlier chunk did), so that it overlaps the copies of earlier results. */
gpv_status gpv_u_nzentries(gpv_handle* h, const char* covType, const double* covparms, int ncovparms,
                           const double* nuggets, const double* nuggets_obsord, int64_t n,
                           double* Lentries, double* Zentries, int64_t* nfail, int64_t* first_fail);

/* Same values, delivered directly in the order createU.R:158-160 builds: for each row of the
 * shard its n0 values (farthest neighbour first, self last), rows concatenated; if Zentries_tail
 * != 0 the 2n Z values follow.  out must hold gpv_packed_len(h) (+2n) doubles. */
int64_t gpv_packed_len(const gpv_handle* h);
int64_t gpv_nuggets_read(const gpv_handle* h);
gpv_status gpv_u_values_packed(gpv_handle* h, const char* covType, const double* covparms,
                               int ncovparms, const double* nuggets, const double* nuggets_obsord,
                               int64_t n, int zentries_tail, double* out, int64_t* nfail,
                               int64_t* first_fail);

/* ---- `covmodel` given as a matrix: U_NZentries_mat (src/U_NZentries.cpp:126-197, createU.R:149-151)
 * covVals: the user's Nlocs x Nlocs covariance of the ordered locations (column-major, host).  Per row
 * covmat = covVals(inds, inds): like the reference, NO nugget is added and revCond is not read.  Same
 * outputs, orders and failure semantics as gpv_u_nzentries / gpv_u_values_packed.  Not a throughput
 * path (the matrix is uploaded per call and caps Nlocs at a few ten thousand). */
gpv_status gpv_u_nzentries_mat(gpv_handle* h, const double* covVals, const double* nuggets_obsord, int64_t n,
                               double* Lentries, double* Zentries, int64_t* nfail, int64_t* first_fail);
gpv_status gpv_u_values_packed_mat(gpv_handle* h, const double* covVals, const double* nuggets_obsord,
                                   int64_t n, int zentries_tail, double* out, int64_t* nfail,
                                   int64_t* first_fail);

/* ---- sparsity arrays and compressed-column output --------------------------
 * Need `obs` at create time.  For a handle that covers all rows these are the arguments of
 * Matrix::sparseMatrix at R/createU.R:161 and the slots of the dgCMatrix it returns; for a shard
 * they are the slices that belong to the shard's rows (= a contiguous range of U's columns).
 *
 * gpv_csc_dims: ncols = U columns covered (rows of the shard + observed rows of the shard), nnz = their
 *   nonzeros (= gpv_packed_len + 2 * observed rows of the shard), size = N + n (U_sparsity.R:12).
 * gpv_u_sparsity: colindices[nnz], rowpointers[nnz], 1-based, in the value order of
 *   gpv_u_values_packed(.., zentries_tail = 1): the loops of R/U_sparsity.R:36-73, bit-identical
 *   (`colindices` is what createU passes as i =, `rowpointers` as j =).
 * gpv_u_csc_pattern: colptr[ncols + 1] (0-based, relative to the shard's first nonzero) and
 *   rowidx[nnz] (0-based U rows, ascending inside a column): dgCMatrix@p and @i.
 * gpv_u_values_csc: x[nnz] in that order (dgCMatrix@x): the kernel, the mask/transpose/concatenate
 *   of createU.R:158-160 and the triplet sort of sparseMatrix in one call.
 * A conditioning set that names the same U row twice (sparseMatrix would sum the two values) makes
 * the last two return GPV_ERR_UNSUPPORTED; gpv_u_sparsity + gpv_u_values_packed still apply. */
gpv_status gpv_csc_dims(gpv_handle* h, int64_t* ncols, int64_t* nnz, int64_t* size);
gpv_status gpv_u_sparsity(gpv_handle* h, int32_t* colindices, int32_t* rowpointers);
gpv_status gpv_u_csc_pattern(gpv_handle* h, int32_t* colptr, int32_t* rowidx);
gpv_status gpv_u_values_csc(gpv_handle* h, const char* covType, const double* covparms, int ncovparms,
                            const double* nuggets, const double* nuggets_obsord, int64_t n, double* x,
                            int64_t* nfail, int64_t* first_fail);

/* ---- fused U + likelihood numerator (vecchia_likelihood.R:74-76), no U materialisation -------
 * zord[n] = z[ord.z].  skip_rows: the first skip_rows locations are `zy` dummies whose latent
 * column createU.R:166-171 drops (0 otherwise).  out[0] = quadform.num contribution,
 * out[1] = logdet.num contribution, out[2] = number of failed rows, all restricted to this
 * handle's row shard; the obs terms (sum z_i^2/tau_i, sum log tau_i) are added by the handle
 * whose shard starts at row 0 (include_obs_terms != 0 forces/disables: -1 auto, 0 no, 1 yes). */
gpv_status gpv_loglik_numerator(gpv_handle* h, const char* covType, const double* covparms,
                                int ncovparms, const double* nuggets, const double* nuggets_obsord,
                                const double* zord, int64_t n, int64_t skip_rows,
                                int include_obs_terms, double out[3]);

/* ---- whole log-likelihood on the GPU for standard Vecchia (`cond.yz = "z"`, every location observed)
 * With pure `z` conditioning U_y U_y^T is diagonal, so the denominator of vecchia_likelihood_U
 * (R/vecchia_likelihood.R:85-91: U2V, sparse Cholesky, triangular solve) reduces to per-row closed
 * forms W_kk = x_kk^2 + 1/tau_k and z2_k = x_kk q_k - z_k/tau_k, accumulated by the same fused kernel.
 * out[0] = log-likelihood (:95-96; meaningful when the handle holds every row), out[1] = quadform.num,
 * out[2] = logdet.num, out[3] = quadform.denom, out[4] = logdet.denom, out[5] = failed rows; parts
 * 1..4 are restricted to the handle's row shard (sum them over ranks).  GPV_ERR_UNSUPPORTED for any
 * other layout (SGV, y, zy, prediction): use gpv_loglik_numerator + the reference's denominator. */
gpv_status gpv_loglik_z(gpv_handle* h, const char* covType, const double* covparms, int ncovparms,
                        const double* nuggets, const double* nuggets_obsord, const double* zord,
                        int64_t n, int include_obs_terms, double out[6]);

/* Data residency for the estimation loop (vecchia_estimate, R/vecchia_wrappers.R:72-93: z is fixed and the
 * nugget is one scalar parameter).  In gpv_loglik_numerator / gpv_loglik_z, zord == NULL reuses the data of
 * the previous likelihood call on the handle, and nuggets == nuggets_obsord == NULL reuses its nuggets or
 * those of gpv_set_scalar_nugget, which builds nuggets.all.ord / nuggets.ord of a scalar nugget on the
 * device (createU.R:70-78: the nugget at observed locations, 0 elsewhere).  A call then moves ~30 bytes
 * each way.  GPV_ERR_ARG if nothing is resident.  gpv_u_nzentries / gpv_u_values_packed / gpv_u_values_csc
 * accept nuggets == nuggets_obsord == NULL the same way (a scalar nugget then costs no per-call upload of
 * 8 bytes per location); a U-values call that is GIVEN nugget vectors replaces the resident ones and drops
 * the flag. */
gpv_status gpv_set_scalar_nugget(gpv_handle* h, double nugget);

/* ---- device-resident variants (inputs/outputs already in HBM; asynchronous on `stream`) ------
 * d_nuggets[Nlocs] device pointer.  d_out: row-major (rows x p) when packed == 0, packed order
 * otherwise.  stream: a cudaStream_t cast to void* (NULL = the handle's own stream).
 * d_loglik (may be NULL): 5 doubles {quadform.num rows part, logdet.num rows part, nfail,
 * quadform.denom, logdet.denom (pure `z` layouts only, else 0)}; requires
 * d_zord (n doubles, device) and obs at create time.  d_out may be NULL when only the
 * likelihood is wanted. */
gpv_status gpv_u_dev(gpv_handle* h, const char* covType, const double* covparms, int ncovparms,
                     const double* d_nuggets, double* d_out, int packed, const double* d_zord,
                     int64_t skip_rows, double* d_loglik, void* stream);

/* Milliseconds of the last U kernel launch of this handle, from CUDA events recorded on the
 * launching stream around that launch (synchronises on the stop event). */
gpv_status gpv_last_kernel_ms(gpv_handle* h, float* ms);
/* Sum of the set-kernel durations (CUDA events on the launching stream, one pair per launch, ring
 * of 128) since the last reset: *count launches, *total_ms milliseconds. */
gpv_status gpv_kernel_time_stats(gpv_handle* h, int reset, int64_t* count, double* total_ms);
/* Name of the kernel instantiation the last launch used, e.g. "u_sets<P=32,D=2>". */
const char* gpv_last_kernel_name(const gpv_handle* h);
/* Number of kernels this library launched since load (monotone counter; bench's gpu_launches). */
int64_t gpv_launch_count(void);

/* ---- one process, several GPUs (the form an R session uses) -------------------------------------
 * One handle per device, driven by worker threads inside the library.  Rows are split into
 * contiguous ranges balancing sum n0^3; each device writes its slice of the packed vector directly
 * into `out`; likelihood partial sums are added on the host in device order.  There is no
 * inter-GPU data path (rows are independent), so no collective is involved.  `devices` may repeat
 * an ordinal (several shards on one GPU). */
/* ---- multi-process runs (one process per GPU): the path's one exchange step ---------------------------------
 * Every rank needs ALL per-location nuggets and all of z (a row's neighbours may be any earlier location) but only
 * its own rows of everything else.  gpv_loglik_z_dist takes each rank's SLICE of those vectors, exchanges the
 * slices GPU to GPU over NVLink / NVSwitch (grouped NCCL broadcasts) and combines the likelihood partial sums with
 * one ncclAllReduce, so that every rank returns the complete log-likelihood -- the north-star's "partial sums
 * combined by an NCCL allreduce over NVLink" behind the C ABI.  NCCL is loaded at run time (libnccl.so.2); without
 * it these entry points return GPV_ERR_UNSUPPORTED.  Bootstrap as in any NCCL program: rank 0 calls
 * gpv_dist_unique_id, the caller carries the 128 bytes to the other ranks, every rank calls gpv_dist_init on ITS
 * shard handle (gpv_create_shard / a row range).  loc_cuts / obs_cuts: world + 1 entries, identical on all ranks;
 * rank r passes entries [loc_cuts[r], loc_cuts[r+1]) of nuggets.all.ord and [obs_cuts[r], obs_cuts[r+1]) of
 * nuggets.ord / zord.  NULL slices on every rank reuse the vectors of the previous call (estimation loop). */
gpv_status gpv_dist_unique_id(void* id128 /* 128 bytes out */);
gpv_status gpv_dist_init(gpv_handle* h, const void* id128, int rank, int world);
gpv_status gpv_dist_finalize(gpv_handle* h);
gpv_status gpv_loglik_z_dist(gpv_handle* h, const char* covType, const double* covparms, int ncovparms,
                             const double* nuggets_slice, const double* tau_slice, const double* z_slice,
                             const int64_t* loc_cuts, const int64_t* obs_cuts, double out[6]);

typedef struct gpv_multi gpv_multi;
gpv_status gpv_multi_create(gpv_multi** out, int64_t Nlocs, int p, int d, const double* locs,
                            const int32_t* revNNarray, const void* revCondOnLatent,
                            gpv_cond_type cond_type, const int32_t* obs, const int* devices, int ndev);
void gpv_multi_destroy(gpv_multi* m);
int gpv_multi_num_devices(const gpv_multi* m);
int64_t gpv_multi_packed_len(const gpv_multi* m);
void gpv_multi_row_cuts(const gpv_multi* m, int64_t* cuts /* ndev + 1 */);
gpv_status gpv_multi_set_revcond(gpv_multi* m, const void* revCondOnLatent, gpv_cond_type cond_type);
gpv_status gpv_multi_u_values_packed(gpv_multi* m, const char* covType, const double* covparms,
                                     int ncovparms, const double* nuggets, const double* nuggets_obsord,
                                     int64_t n, int zentries_tail, double* out, int64_t* nfail,
                                     int64_t* first_fail);
gpv_status gpv_multi_loglik_numerator(gpv_multi* m, const char* covType, const double* covparms,
                                      int ncovparms, const double* nuggets, const double* nuggets_obsord,
                                      const double* zord, int64_t n, int64_t skip_rows, double out[3]);
gpv_status gpv_multi_loglik_z(gpv_multi* m, const char* covType, const double* covparms, int ncovparms,
                              const double* nuggets, const double* nuggets_obsord, const double* zord,
                              int64_t n, double out[6]);
/* compressed-column output over the devices (see gpv_u_csc_pattern / gpv_u_values_csc): same arrays as a
 * single full-range handle returns, every device filling the slice of its row shard */
gpv_status gpv_multi_csc_dims(gpv_multi* m, int64_t* ncols, int64_t* nnz, int64_t* size);
gpv_status gpv_multi_u_csc_pattern(gpv_multi* m, int32_t* colptr, int32_t* rowidx);
gpv_status gpv_multi_u_values_csc(gpv_multi* m, const char* covType, const double* covparms, int ncovparms,
                                  const double* nuggets, const double* nuggets_obsord, int64_t n, double* x,
                                  int64_t* nfail, int64_t* first_fail);
/* sets the calling thread's gpv_last_error() text (used by the multi-GPU front end) */
void gpv_set_last_error(const char* msg);

/* ---- stateless drop-in: the reference's nine arguments, everything uploaded per call ---------
 * Mirrors U_NZentries(Ncores, n, locs, revNNarray, revCondOnLatent, nuggets, nuggets_obsord,
 * covType, covparms) (src/U_NZentries.cpp:25); Ncores is accepted and ignored. */
gpv_status gpv_U_NZentries(int Ncores, int64_t n, int64_t Nlocs, int p, int d, const double* locs,
                           const int32_t* revNNarray, const void* revCondOnLatent,
                           gpv_cond_type cond_type, const double* nuggets,
                           const double* nuggets_obsord, const char* covType,
                           const double* covparms, int ncovparms, double* Lentries,
                           double* Zentries, int64_t* nfail, int64_t* first_fail, int device);
/* gpv_U_NZentries keeps the handle of its last call and recognises the next call's locs / revNNarray by a 64-bit
 * fingerprint (a different revCond alone is re-uploaded; GPV_STATELESS_CACHE=0: create and destroy per call).
 * This frees the kept handle and its device memory (package unload). */
void gpv_release_cached(void);

/* ---- covariance functions alone (MaternFun / EsqeFun, exported by the reference:
 * R/RcppExports.R:4-17, src/Matern.cpp:24, src/Esqe.cpp:17).  dist/out: len doubles on host. */
gpv_status gpv_MaternFun(const double* dist, int64_t len, const double* covparms, double* out,
                         int device);
gpv_status gpv_EsqeFun(const double* dist, int64_t len, const double* covparms, double* out,
                       int device);

/* ---- measurement helpers ---------------------------------------------------------------------
 * fp64 FMA-pipe peak of `device` from a register-resident DFMA loop (TFLOP/s, 2 flop per FMA)
 * and a device-to-device copy bandwidth (GB/s, read+write bytes), both timed with CUDA events. */
gpv_status gpv_measure_fp64_peak(int device, double* tflops);
gpv_status gpv_measure_copy_bw(int device, double* gbs);

/* ---- ic0 / createUcppM / createUcpp (src/ic0.cpp:43-63, :68-71, :77-92; R/RcppExports.R:53-63) -------
 * The incomplete-Cholesky (MRA) branch of createU (R/createU.R:89-106).  The pattern is compressed sparse
 * row, lower triangular, every row ending on its diagonal: ptrs (N + 1) and inds (nnz) are R numeric
 * vectors holding 0-based integers, exactly what createU.R:90-91 builds.  gpv_ic0 overwrites vals (nnz)
 * with the incomplete-Cholesky values, like the reference (which modifies its argument in place and
 * returns it); gpv_createUcppM is the same call under the name createU uses for a matrix or function
 * covmodel.  gpv_createUcpp first fills vals with MaternFun(|locsord[i,] - locsord[inds[j],]|, covparams)
 * for every stored entry -- on `device`, one thread per entry -- and then runs ic0 on the host (row i reads
 * the finished rows of its columns: a sequential sweep, as in the reference).  locsord: N x d column-major.
 * Unlike the reference (which prints "ERROR" and carries on, :57-58) an entry right of the diagonal, a
 * non-integer or out-of-range index, or a column whose own row is empty is GPV_ERR_ARG and vals is left
 * untouched. */
gpv_status gpv_ic0(int64_t N, const double* ptrs, const double* inds, int64_t nnz, double* vals);
gpv_status gpv_createUcppM(int64_t N, const double* ptrs, const double* inds, int64_t nnz, double* cov_vals);
gpv_status gpv_createUcpp(int64_t N, int d, const double* ptrs, const double* inds, int64_t nnz,
                          const double* locsord, const double* covparams, double* vals, int device);

/* ---- whichCondOnLatent (R/whichCondOnLatent.R:2-27), host code -------------------------------
 * The sparse-general-Vecchia rule (vecchia_specify.R:183-185): NNarray is the un-reversed n x p neighbour
 * array, column-major, 1-based, missing = NA_integer_ (or 0); firstind_pred <= 0 means n + 1.
 * CondOnLatent (n x p, column-major) receives R logicals: 1, 0, NA_integer_.  Sequential over rows like
 * the reference (row k reads the finished rows of its neighbours); runs on the host, no device needed. */
gpv_status gpv_whichCondOnLatent(const int32_t* NNarray, int64_t n, int p, int64_t firstind_pred,
                                 int32_t* CondOnLatent);

/* ---- harness: synthetic inputs at scale (not part of the reference path) -------
 * Ordered m-nearest-neighbour search on the GPU: row i gets (i, its min(m,i) nearest among rows
 * < i, nearest first) -- semantics of GpGp::find_ordered_nn as used at vecchia_specify.R:159 --
 * written as a column-reversed, 1-based, 0-padded revNNarray (U_sparsity.R:32), column-major,
 * for rows [row_begin,row_end).  locs column-major Nlocs x d (d = 2 or 3), host. */
gpv_status gpv_harness_ordered_nn(int64_t Nlocs, int d, int m, const double* locs, int64_t row_begin,
                                  int64_t row_end, int32_t* revNNarray_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* GPVECCHIA_B200_H */
