/*
 * mlmcb200.h -- C ABI of the B200-native MLMC estimation hot path (libmlmcb200.so).
 *
 * The reference (GeoMop/MLMC v1.0.2) is pure Python and has no FFI; its de-facto operator interface for
 * this path is duck-typed Python.  Each entry point below names the reference
 * code it replaces (paths relative to the reference root).  All `double*` / `uint8_t*` data arguments are
 * DEVICE pointers owned by the caller (torch tensors in the Python host layer); `stream` is a
 * `cudaStream_t` passed as `void*`; every call is asynchronous on that stream and keeps no global state.
 * Return value: 0 = OK, negative = error (text via mlmcb200_last_error(), thread-local).
 *
 * Sample layout ("level chunk"): element (sample n, side s, component m) lives at
 *     pairs[n * stride_n + s * stride_side + m * stride_m]        (strides in doubles)
 * side 0 = fine, side 1 = coarse.  The storage row order of the reference, float64[N, 2, M]
 * (mlmc/sample_storage_hdf.py:169-184, mlmc/tool/hdf5.py:311-320), is stride_n = 2M, stride_side = M,
 * stride_m = 1.  Level 0 has no coarse part: pass has_coarse = 0 (the zero coarse row is then never read).
 */
#ifndef MLMCB200_H
#define MLMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLMCB200_ABI_VERSION 2

/* mlmcb200_basis_t.kind */
#define MLMCB200_RAW       0   /* identity, size 1: phi_0(x) = x, no transform (plain estimate_mean of a quantity) */
#define MLMCB200_LEGENDRE  1   /* mlmc/moments.py:174-197  (numpy legvander recurrence)  */
#define MLMCB200_MONOMIAL  2   /* mlmc/moments.py:111-126  (numpy polyvander recurrence) */
#define MLMCB200_FOURIER   3   /* mlmc/moments.py:133-162  (1, cos t, sin t, cos 2t, ...) */

#define MLMCB200_MAX_MOMENTS 256

/* One `mlmc.moments.Moments` object: mlmc/moments.py:10-39 (transform) and :58-73 (clip / linear). */
typedef struct {
    int32_t kind;      /* MLMCB200_* above                                                     */
    int32_t size;      /* number of base functions R                                           */
    int32_t is_log;    /* Moments._is_log : value <- ln(value) before the affine map           */
    int32_t is_clip;   /* Moments._is_clip (safe_eval): t outside [ref_lo, ref_hi] -> NaN      */
    double  shift;     /* Moments._linear_shift                                                */
    double  scale;     /* Moments._linear_scale                                                */
    double  ref_lo;    /* Moments.ref_domain[0]                                                */
    double  ref_hi;    /* Moments.ref_domain[1]                                                */
} mlmcb200_basis_t;

int         mlmcb200_abi_version(void);
const char* mlmcb200_last_error(void);

/* Number of SMs of the current device (grid sizing is done inside the library; exported for reports). */
int mlmcb200_sm_count(void);

/*
 * Moments.eval_all / __call__  (mlmc/moments.py:75-93, :122-126, :145-162, :195-197; TransformedMoments :256-259)
 * x[n] -> out[n][n_out].  `matrix` == NULL: out[:, r] = phi_r(x), n_out <= basis->size functions are written.
 * `matrix` != NULL ([n_rows][basis->size], row-major, device): out = Phi @ matrix^T truncated to n_out <= n_rows.
 * Legendre / Monomial tables are bit-identical to numpy's legvander / polyvander.
 */
int mlmcb200_basis_eval(const mlmcb200_basis_t* basis, const double* x, int64_t n,
                        const double* matrix, int32_t n_rows, int32_t n_out,
                        double* out, void* stream);

/*
 * mask_nan_samples for vector quantities (mlmc/quantity/quantity_estimate.py:6-14 applied to the moments of
 * all M components): valid[n] = 1 iff every component of the fine (and, if has_coarse, coarse) row maps to a
 * non-NaN moment vector.  Scalar quantities (M == 1) do not need it: the accumulate calls test inline.
 */
int mlmcb200_sample_mask(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                         int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                         uint8_t* valid, void* stream);

/*
 * Level accumulator of the fused moments + mean/variance reduction.  Replaces the chunk loop of
 * estimate_mean over a `moments` quantity: mlmc/quantity/quantity_estimate.py:43-65 with the operation
 * of :105-110 and the difference of :59-62.  Layout (doubles):
 *     acc[0] = n_samples (valid), acc[1] = n_rm_samples (masked), acc[2 + k] = sum_n d_k,
 *     acc[2 + K + k] = sum_n d_k^2,   K = n_comp * size,  k = m * size + r  (mom_at_bottom order)
 * where d = phi_r(fine_m) - phi_r(coarse_m) (level 0: phi_r(fine_m)).  The call ADDS the chunk into acc,
 * so a level may be streamed in any number of chunks; zero acc first.
 * `valid` (from mlmcb200_sample_mask) is required when n_comp > 128; for 1 < n_comp <= 128 it may be NULL: the kernel
 * then combines the components' domain tests per sample itself (and size <= 112).  NULL when n_comp == 1.
 * `workspace` must hold mlmcb200_moments_workspace_bytes(...) bytes.
 * Limit: basis->size <= 226 for n_comp == 1, <= 113 otherwise (per-thread accumulator columns in shared memory).
 */
int64_t mlmcb200_moments_workspace_bytes(int32_t size, int32_t n_comp);
int mlmcb200_moments_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                                int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                const uint8_t* valid, double* acc, void* workspace, int64_t workspace_bytes,
                                void* stream);

/*
 * The same when the caller needs the SUMS only (acc[2 + k]; e.g. the moment sums behind the linearised covariance means,
 * mlmcb200_level_sums_transform): the sums of squares acc[2 + K + k] are then unspecified (they may or may not be
 * updated).  Scalar quantities in storage order take a kernel variant without the squares (6 instead of 7.25 FP64
 * instructions per sample-moment, half the shared memory per thread).
 */
int mlmcb200_moments_accumulate_sums(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                                     int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                     const uint8_t* valid, double* acc, void* workspace, int64_t workspace_bytes,
                                     void* stream);

/*
 * Bootstrap re-sampling of one level on the device.  Replaces the replicate loop of Estimate.est_bootstrap
 * (mlmc/estimator.py:171-218) over Quantity.subsample / pick_samples (mlmc/quantity/quantity.py:307-364: `size`
 * rows drawn WITH replacement from the chunk) followed by estimate_mean of the moments: replicate b, b < n_rep,
 * accumulates the rows  pairs[idx[b * n_draws + i]], i < n_draws  (idx: device, row numbers < n_rows, drawn by
 * the caller) exactly as mlmcb200_moments_accumulate would accumulate that gathered chunk, into
 * acc + b * acc_rep_stride (same layout; ADDS).  `valid` is indexed by storage row.  All replicates of the level
 * run in ONE launch (grid.z = replicate); rows are re-read through L2, never materialised per replicate.
 */
/*
 * Row numbers for mlmcb200_moments_accumulate_resampled, drawn on the device: idx[b * n_draws + j], b < n_rep, is
 * uniform on [0, n_rows) WITH replacement (RNG.choice of Quantity.pick_samples, mlmc/quantity/quantity.py:318-319).
 * Counter-based Philox4x32-10: key (seed, stream_id), counter (draw, rep_offset + b) -- replicate rep_offset + b gets
 * the same rows whatever the grouping of the replicates into calls (or their sharding over ranks).
 * n_blocks > 1 orders each replicate's draws by row block (block p = rows [p n_rows / P, (p+1) n_rows / P)) so that
 * concurrent CTAs gather from an L2-sized window: block_cum (device, [n_rep][n_blocks + 1], block_cum[b][0] = 0,
 * block_cum[b][P] = n_draws) are the caller's cumulative MULTINOMIAL block counts -- the draws are then exactly
 * i.i.d. uniform rows, listed in block order (the level sums do not depend on the order).
 */
int mlmcb200_resample_indices(uint64_t seed, uint64_t stream_id, int64_t n_rows, int64_t n_draws, int32_t n_rep,
                              int32_t rep_offset, int32_t n_blocks, const int64_t* block_cum, int32_t* idx,
                              void* stream);
int64_t mlmcb200_moments_resampled_workspace_bytes(int32_t size, int32_t n_comp, int32_t n_rep);
int mlmcb200_moments_accumulate_resampled(const mlmcb200_basis_t* basis, const double* pairs, int64_t n_rows,
                                          int32_t n_comp, int64_t stride_n, int64_t stride_side, int64_t stride_m,
                                          int32_t has_coarse, const uint8_t* valid, const int32_t* idx,
                                          int64_t n_draws, int32_t n_rep, double* acc, int64_t acc_rep_stride,
                                          void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Bootstrap re-sampling of one level in ONE pass over its rows (scalar quantity; Legendre, Monomial or Fourier basis of at most
 * mlmcb200_moments_weighted_max_size() moments; bases of more than 51 moments take ceil((2 + 2 R) / 104) passes, one
 * per group of 104 result columns): the same replicate loop as above (mlmc/estimator.py:171-218,
 * mlmc/quantity/quantity.py:307-322), written as the weighted sums
 *     acc[b][2 + r] += sum_i w_bi d_r(i),   acc[b][2 + R + r] += sum_i w_bi d_r(i)^2,   acc[b][0 / 1] += sum_i w_bi ok_i / rm_i
 * with w_bi = how many times replicate b drew row i: the dense product W^T [ok | rm | D | D.D] on FP64 tensor-core
 * tiles (DMMA m8n8k4).  The basis recurrence of a row runs once for ALL replicates and the rows are read once, in
 * storage order.  The better choice when a replicate draws about as many rows as the level holds (the usual bootstrap);
 * few draws from many rows stay with mlmcb200_moments_accumulate_resampled.
 *
 * mlmcb200_resample_counts: counts[b * counts_stride + i] (device bytes) = multiplicity of row i among the draws
 * block_cum[b][0] .. block_cum[b][n_blocks] of replicate rep_offset + b -- exactly the draws mlmcb200_resample_indices
 * lists for the same (seed, stream_id, n_rows, n_blocks, block_cum); block_cum is required here (n_blocks = 1:
 * {0, n_draws}).  Row blocks hold at most mlmcb200_resample_counts_block_rows() rows (byte counters of a block live in
 * shared memory); max_draws = the largest block_cum[b][n_blocks], at most 8 n_rows (byte counters).
 * mlmcb200_moments_accumulate_weighted: pairs = the level's rows in storage order (has_coarse: (fine, coarse) pairs,
 * stride_n = 2; level 0: stride_n doubles apart); counts rows 16-byte aligned, counts_stride % 16 == 0.  ADDS into
 * acc + b * acc_rep_stride (layout of mlmcb200_moments_accumulate).
 */
int32_t mlmcb200_resample_counts_block_rows(void);
int mlmcb200_resample_counts(uint64_t seed, uint64_t stream_id, int64_t n_rows, int64_t max_draws, int32_t n_rep,
                             int32_t rep_offset, int32_t n_blocks, const int64_t* block_cum, uint8_t* counts,
                             int64_t counts_stride, void* stream);
int32_t mlmcb200_moments_weighted_max_size(void);
int64_t mlmcb200_moments_weighted_workspace_bytes(int64_t n_rows, int32_t n_rep, int32_t size);
int mlmcb200_moments_accumulate_weighted(const mlmcb200_basis_t* basis, const double* pairs, int64_t n_rows,
                                         int64_t stride_n, int32_t has_coarse, const uint8_t* counts,
                                         int64_t counts_stride, int32_t n_rep, double* acc, int64_t acc_rep_stride,
                                         void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Level accumulator of the moment-covariance estimate (scalar quantity).  Replaces estimate_mean over a
 * `covariance` quantity: mlmc/quantity/quantity_estimate.py:131-147 + :43-65.  With R = basis->size:
 *     acc[0] = n_samples, acc[1] = n_rm_samples,
 *     acc[2 + i*R + j]       = sum_n d_ij,   d_ij = phi_i(f) phi_j(f) - phi_i(c) phi_j(c)   (level 0: fine only)
 *     acc[2 + R*R + i*R + j] = sum_n d_ij^2  (only when want_var != 0; otherwise untouched)
 * Both are dense contractions Phi^T Phi and run on FP64 tensor-core tiles (DMMA m8n8k4).
 * mode: 0 = covariance (above); 1 = Gram of the differences, sum_n d_i d_j with d = phi(f) - phi(c),
 * which is what TransformedMoments needs for its variances (var(L d) = L G L^T), want_var ignored.
 * Limit: basis->size <= 104 (13 blocks of 8 moments in the shared-memory tile); larger sizes return an error.
 */
int64_t mlmcb200_gram_workspace_bytes(int32_t size);
int mlmcb200_gram_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                             int64_t stride_n, int64_t stride_side, int32_t has_coarse,
                             int32_t mode, int32_t want_var, double* acc,
                             void* workspace, int64_t workspace_bytes, void* stream);

/*
 * The same for a quantity of n_comp components (the reference's covariance quantity handles any shape,
 * mlmc/quantity/quantity_estimate.py:131-147: outer products per component, cov_at_bottom order): component m of
 * sample n at pairs[n * stride_n + s * stride_side + m * stride_m]; K = n_comp * R * R and
 *     acc[2 + m*R*R + i*R + j] = sum_n d_ij (component m),   acc[2 + K + m*R*R + i*R + j] = sum_n d_ij^2.
 * `valid` (mlmcb200_sample_mask, shared by all components: mask_nan_samples drops a sample when ANY component leaves
 * the domain) is required for n_comp > 1, NULL for n_comp == 1.  One launch covers all components (grid.y), the SMs
 * are shared out between them.
 */
int64_t mlmcb200_gram_workspace_bytes_comp(int32_t size, int32_t n_comp);
int mlmcb200_gram_accumulate_comp(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                                  int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                  const uint8_t* valid, int32_t mode, int32_t want_var, double* acc,
                                  void* workspace, int64_t workspace_bytes, void* stream);

/*
 * l_means / l_vars of estimate_mean (mlmc/quantity/quantity_estimate.py:70-77) and QuantityMean.mean / .var
 * (mlmc/quantity/quantity.py:588-593) for n_levels accumulators of K sums each, laid out back to back
 * with a stride of acc_stride doubles:  mean_l = s/n,  var_l = (sq - s^2/n)/(n-1)  (n <= 1 -> +inf).
 * out: l_means[L][K], l_vars[L][K], mean[K], var[K] (any may be NULL).
 */
int mlmcb200_finalize_levels(const double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                             double* l_means, double* l_vars, double* mean, double* var, void* stream);

/*
 * One whole estimate_mean over a `moments` quantity in ONE call (mlmc/quantity/quantity_estimate.py:22-80 with the
 * operation of :105-110): zero the accumulators, mlmcb200_moments_accumulate of every level (levels[l]: host array of
 * descriptors of device rows, one chunk per level), mlmcb200_finalize_levels, and -- if host_out != NULL (pinned host
 * memory) -- the copy of the packed result to the host followed by a stream synchronisation.  For callers whose levels
 * are resident in device memory: the ~10 separate host calls of the step-by-step route cost more than the kernels of a
 * small estimate (cfg1: 1e5 samples).  out / host_out (doubles):
 *     [l_means (L*K) | l_vars (L*K) | mean (K) | var (K) | n_samples, n_rm_samples per level (2 L)],  K = n_comp * size.
 */
typedef struct {
    const double*  pairs;        /* device rows of the level (layout above) */
    int64_t        n;            /* samples */
    int64_t        stride_n, stride_side, stride_m;
    int32_t        has_coarse;
    int32_t        reserved;
    const uint8_t* valid;        /* mlmcb200_sample_mask result for n_comp > 128, else NULL */
} mlmcb200_level_t;
int mlmcb200_estimate_moments_levels(const mlmcb200_basis_t* basis, const mlmcb200_level_t* levels, int32_t n_levels,
                                     int32_t n_comp, double* acc, int64_t acc_stride, double* out, double* host_out,
                                     void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Covariance level sums from moment level sums.  The reference forms per-sample outer products phi_i phi_j
 * (mlmc/quantity/quantity_estimate.py:131-147) and sums them; products of Legendre / monomial / trigonometric
 * functions are exact linear combinations of a longer basis of the same family (phi_i phi_j = sum_k C[ij][k] phi_k,
 * k < K0), and the NaN mask depends on the domain only, so per level
 *     sum_n d_ij = sum_k C[ij][k] * (sum_n d_k)
 * with the sums of mlmcb200_moments_accumulate over the K0-function basis.  This call applies such a map:
 *     acc_out[l][2 + m K1 + o] = sum_k mat_t[k K1 + o] * acc_in[l][2 + m K0 + k],   o < K1, m < n_comp, l < n_levels
 * (mat_t: device, [K0][K1] row-major), copies the two counts and fills the sums of squares of acc_out with NaN
 * (mlmcb200_finalize_levels then yields NaN variances: entry variances need mlmcb200_gram_accumulate).
 * Also the map TransformedMoments (mlmc/moments.py:256-259) applies under the sums.
 */
int mlmcb200_level_sums_transform(const double* acc_in, int64_t in_stride, int32_t n_levels, int32_t K0,
                                  int32_t n_comp, const double* mat_t, int32_t K1, double* acc_out,
                                  int64_t out_stride, void* stream);

/*
 * The same for n_batch accumulator sets acc_batch_stride doubles apart (the replicates of est_bootstrap,
 * mlmc/estimator.py:185-193): out[b] = [l_means (L*K) | l_vars (L*K) | mean (K) | var (K)], packed per entry.
 */
int mlmcb200_finalize_levels_batched(const double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                                     int32_t n_batch, int64_t acc_batch_stride, double* out, void* stream);

/*
 * Multi-GPU: finalize fused with the cross-rank SUM of the level accumulators over NVLink peer memory (one launch,
 * no NCCL call, no host synchronisation).  The reference has no distributed estimation; this replaces the
 * "all-reduce, then mlmcb200_finalize_levels" pair of the sample-sharded estimate.
 *   peer_buffer_bytes / peer_alloc : every rank allocates one exchange buffer (cudaMalloc, zeroed) and gets its
 *                                    64-byte cudaIpc handle, which the host side exchanges between the ranks
 *   peer_open / peer_close         : map / unmap another rank's buffer in this process
 *   allreduce_finalize_levels      : acc [n_levels][acc_stride] of this rank is added over all ranks IN RANK ORDER
 *                                    (bit-identical sums everywhere) and written back to acc; l_means / l_vars /
 *                                    mean / var as mlmcb200_finalize_levels.  peer_buffers: DEVICE array [world] of the
 *                                    ranks' buffers as mapped here (own buffer at index rank).  Collective: every
 *                                    rank must call it the same number of times.  Waits at most
 *                                    MLMCB200_PEER_TIMEOUT_MS (environment, default 30 000) for the peers; a rank that
 *                                    gives up marks the epoch in every peer, so late ranks fail it as well.  A failed call
 *                                    leaves acc (the LOCAL sums) untouched, writes NaN results, status[0] = 1 (device,
 *                                    may be NULL; 0 on success) and sets the sticky error word (mlmcb200_peer_error):
 *                                    the caller then falls back to an all-reduce + mlmcb200_finalize_levels.
 */
int64_t mlmcb200_peer_buffer_bytes(int32_t world, int64_t slot_doubles);
int mlmcb200_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
int mlmcb200_peer_open(const unsigned char* handle64, void** dev_ptr);
int mlmcb200_peer_close(void* dev_ptr);
int mlmcb200_peer_free(void* dev_ptr);
int mlmcb200_peer_error(const void* own_buffer, int32_t world, int64_t slot_doubles, int32_t* error);
int mlmcb200_allreduce_finalize_levels(double* acc, int64_t acc_stride, int32_t n_levels, int64_t K, int32_t rank,
                                       int32_t world, void* const* peer_buffers, int64_t slot_doubles,
                                       double* l_means, double* l_vars, double* mean, double* var, double* status,
                                       void* stream);

/*
 * Order statistics for Estimate.estimate_domain (mlmc/estimator.py:275-302): the reference takes
 * np.percentile(fine, [100 q, 100 (1 - q)]) of a level's fine samples.  For every fraction f in frac[0..n_frac)
 * (host array, values in [0, 1]) this finds, among the non-NaN entries x[i * stride], i < n (device), the two
 * neighbouring order statistics numpy interpolates between: with pos = f * (n_valid - 1),
 *     out[2 k] = sorted[floor(pos)],  out[2 k + 1] = sorted[min(floor(pos) + 1, n_valid - 1)],  out[2 n_frac] = n_valid
 * (device, doubles).  Exact radix selection on the 64-bit keys (six histogram passes), no sort and no copy of x.
 */
int64_t mlmcb200_percentile_workspace_bytes(int32_t n_frac);
int mlmcb200_percentile_stats(const double* x, int64_t n, int64_t stride, const double* frac, int32_t n_frac,
                              double* out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Max-entropy functional pieces on a fixed node set (mlmc/tool/simple_distribution.py:254-327):
 *     rho_q = exp(clip(-phi_q . lam_scaled, -200, 200)),   lam_scaled = lambda / sigma
 *     out[0]               = sum_q w_q rho_q                       (integral term of _calculate_functional)
 *     out[1 + i]           = sum_q w_q rho_q phi_qi                (integral term of _calculate_gradient, before /sigma)
 *     out[1 + R + i*R + j] = sum_q w_q rho_q phi_qi phi_qj         (_calculate_jacobian_matrix, before /sigma_i sigma_j)
 * phi is [Q][ld] row-major with R <= ld.  what: bit 0 = F, bit 1 = g, bit 2 = H (unrequested parts are left untouched).
 */
int64_t mlmcb200_maxent_workspace_bytes(int64_t n_nodes, int32_t size);
int mlmcb200_maxent_fgh(const double* phi, int64_t ld, const double* w, const double* lam_scaled,
                        int64_t n_nodes, int32_t size, int32_t what, double* out,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * SimpleDistribution.density (mlmc/tool/simple_distribution.py:96-105) for values x[n] (device):
 *     out[k] = exp(clip(-sum_{i < n_coef} coef[i] phi_i(x[k]), -200, 200)),   coef = lambda / sigma  (device, n_coef <= size)
 * phi as mlmcb200_basis_eval (same operations); a value outside a clipped domain gives NaN, as in the reference.
 * For a TransformedMoments basis (phi' = L phi) pass coef = L[:n]^T (lambda / sigma) and the base functions.
 */
int mlmcb200_density_eval(const mlmcb200_basis_t* basis, const double* x, int64_t n, const double* coef,
                          int32_t n_coef, double* out, void* stream);

/* FP64 pipe micro-benchmarks used by bench.py for the roofline denominators (not on the data path).
 * kind & 15: 0 = DFMA with two loop-invariant operands, 1 = DMMA m8n8k4, 2 = DFMA with three distinct register
 * operands: sustained throughput in FLOP/s (FMA = 2), CUDA events; 3 = latency of a dependent DFMA in SM cycles.
 * kind >> 4: resident warps per SM to run with (0 = full occupancy). */
int mlmcb200_fp64_peak(int32_t kind, double* flops_per_s, void* stream);

/*
 * Host helper of the staged feed (SampleStorageHDF / NpyStorage -> pinned staging buffer -> device; the read side of
 * mlmc/sample_storage_hdf.py:169-184, mlmc/tool/hdf5.py:353-376): n_rows rows of keep_bytes bytes, src_pitch bytes apart
 * in src (a file mapping), packed into dst by n_threads host threads.  keep_bytes < src_pitch drops the tail of every
 * row (level 0 without its stored zero coarse row).  No device work.
 */
int mlmcb200_host_copy_rows(void* dst, const void* src, int64_t n_rows, int64_t keep_bytes, int64_t src_pitch,
                            int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* MLMCB200_H */
