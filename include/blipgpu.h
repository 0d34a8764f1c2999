/*
 * blipgpu.h -- C ABI of libblipgpu.so: the B200 (sm_100a) implementation of
 * pyblip's multi-chain Gibbs samplers and candidate-group PIP counting.
 *
 * This is the drop-in boundary.  Every entry point names the reference
 * interface it replaces (paths relative to amspector100/pyblip v0.4.2).  The
 * reference reaches its hot path through Python -> Cython calls
 * (`apply_pool(fn, ...)` in pyblip/utilities.py:63-139 mapping one Cython call
 * per chain); a maintainer binds these symbols with ctypes instead (see
 * INTEGRATION.md).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a BLIPGPU_ERR_* code otherwise;
 *     blipgpu_last_error() returns a thread-local message for the last failure;
 *   - host pointers are borrowed for the duration of the call only;
 *   - handles are opaque and freed by the matching *_free;
 *   - "device pointer" arguments must belong to the context's device.
 *
 * Environment (diagnostics; none changes a result)
 *   BLIPGPU_LAUNCH_LOG=1             every sweep-kernel launch with its CUDA-event time and a host time stamp,
 *                                    allocation times and the stages of samples_csr on stderr
 *   BLIPGPU_NO_POOL=1                device buffers from cudaMalloc / cudaFree instead of the library's cached,
 *                                    stream-ordered blocks
 *   BLIPGPU_DEVICE_CACHE_MB=x        idle device blocks kept per stream for the next call (default 65536)
 *   BLIPGPU_HOST_CACHE_MB=x          idle page-locked host blocks kept for the next result (default 8192)
 *   BLIPGPU_NO_PINNED_OUT=1          (host side, pyblip_b200/_lib.py) results in plain NumPy arrays
 *   BLIPGPU_DIST_DEBUG=1             (host side, pyblip_b200/dist.py) timing of every count merge on stderr
 *   BLIPGPU_NO_L2_PIN=1              residual-form sweeps: do not pin the design matrix in L2
 *   BLIPGPU_NO_WS=1                  gram form: never pick the role-specialised few-chains kernels
 *   BLIPGPU_WS_CLUSTER=1             gram form: AUTO picks the 2-CTA cluster placement when chains <= SMs/2
 *   BLIPGPU_WB=1|2                   gram form: AUTO picks the batched-decision kernel (2: its cluster placement)
 *   BLIPGPU_WS2=1 / BLIPGPU_WS2C=1   gram form: AUTO picks the two-position kernel (one SM / 2-CTA cluster)
 *   BLIPGPU_STREAM=1                 gram form: AUTO hands stationary launches over to the column-stream kernel
 *   BLIPGPU_STREAM_MEAN=x            ... when the previous launch saw at most x changes per chain-sweep (default 90)
 *   BLIPGPU_PAIRS_TC=0|1             count_pairs: forbid / force the tensor-core kernel (default: from 512 columns up)
 *   BLIPGPU_COUNT_SEQ_TILES=1        count_sequential: the round-1 tile kernel (A/B timing, max_size <= 64)
 *   BLIPGPU_NPRIOR_WORKSPACE_MB=x    budget of the neuronized prior's joint-draw workspace
 *                                    (default: half of the free device memory; chains run in
 *                                    waves when all of them do not fit)
 */
#ifndef BLIPGPU_H
#define BLIPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLIPGPU_ABI_VERSION 1

#define BLIPGPU_OK 0
#define BLIPGPU_ERR_CUDA 1
#define BLIPGPU_ERR_ARG 2
#define BLIPGPU_ERR_NOMEM 3
#define BLIPGPU_ERR_NUMERIC 4      /* reference raises RuntimeError / ValueError */
#define BLIPGPU_ERR_UNSUPPORTED 5

typedef struct blipgpu_ctx blipgpu_ctx;         /* one device + its streams        */
typedef struct blipgpu_design blipgpu_design;   /* X resident in HBM (+ Gram)      */
typedef struct blipgpu_samples blipgpu_samples; /* posterior samples kept on device */
typedef struct blipgpu_bits blipgpu_bits;       /* bit-packed indicator matrix      */

int blipgpu_abi_version(void);
/* sizeof() of the ABI structs as compiled, for binding self-checks:
 * 0 spikeslab_config, 1 streams, 2 sample_options, 3 samples_info, 4 multi_streams,
 * 5 nprior_config, 6 nprior_streams. */
int blipgpu_sizeof(int which);
const char* blipgpu_last_error(void);
int blipgpu_device_count(int* count);

/* ---- context ----------------------------------------------------------- */
int blipgpu_ctx_create(int device, blipgpu_ctx** out);
int blipgpu_ctx_destroy(blipgpu_ctx* ctx);
int blipgpu_ctx_synchronize(blipgpu_ctx* ctx);
/* Elapsed device milliseconds (CUDA events on the context's stream) spent in
 * sweep kernels / counting kernels since the last reset; used by bench.py. */
int blipgpu_ctx_timing(blipgpu_ctx* ctx, double* sweep_ms, int64_t* sweep_launches,
                       double* other_ms, int64_t* other_launches, int reset);
/* Work counters since the last reset: kernels launched by this library on the
 * context, and bytes it copied host->device / device->host. */
int blipgpu_ctx_counters(blipgpu_ctx* ctx, int64_t* kernel_launches, int64_t* h2d_bytes,
                         int64_t* d2h_bytes, int reset);
/* CUDA-event stopwatch on the context's stream (slot 0..7): bench.py brackets its
 * timed regions with these because torch.cuda.Event only sees torch's stream. */
int blipgpu_ctx_event_record(blipgpu_ctx* ctx, int slot);
int blipgpu_ctx_event_elapsed_ms(blipgpu_ctx* ctx, int slot_begin, int slot_end, double* ms);

/* ---- page-locked host blocks for results ---------------------------------
 * The reference returns fresh NumPy arrays (`np.concatenate([...])`,
 * pyblip/linear/linear.py:165-168).  A result written into page-locked memory is
 * one DMA; into a fresh pageable array it is a bounce copy plus a page fault per
 * page.  The host side therefore takes the arrays it returns from here and hands
 * them back when the array is collected; blocks are recycled process-wide (up to
 * BLIPGPU_HOST_CACHE_MB, default 8192, stay pinned while idle).  Every entry point
 * that writes to host memory accepts either kind of destination. */
int blipgpu_host_alloc(size_t bytes, void** out);
int blipgpu_host_free(void* p);

/* ---- design matrix ----------------------------------------------------- */
#define BLIPGPU_DESIGN_DENSE 0        /* X given, column-major copy kept in HBM   */
#define BLIPGPU_DESIGN_CHANGEPOINT 1  /* X[i][j] = 1{i >= j}, never materialised
                                         (pyblip/changepoint.py:33-35)           */
/* Replaces `XT = np.ascontiguousarray(X.T)` and `Xl2 = (X**2).sum(0)`
 * (pyblip/linear/_linear.pyx:73-74).  X is n x p row-major (C-contiguous) float64
 * on the HOST. */
int blipgpu_design_upload(blipgpu_ctx* ctx, const double* X, int64_t n, int64_t p,
                          blipgpu_design** out);
int blipgpu_design_changepoint(blipgpu_ctx* ctx, int64_t T, blipgpu_design** out);
/* Builds the p x p Gram matrix X^T X on the device (fp64 tensor cores); replaces
 * `np.dot(X.T, X)` (pyblip/linear/_linear_multi.pyx:379).  Idempotent. */
int blipgpu_design_build_gram(blipgpu_design* d);
/* Debug / test access: copy the Gram block rows [r0, r1) x all columns to host. */
int blipgpu_design_get_gram(blipgpu_design* d, int64_t r0, int64_t r1, double* out);
int blipgpu_design_get_xl2(blipgpu_design* d, double* out);
int blipgpu_design_free(blipgpu_design* d);

/* ---- spike-and-slab sampler -------------------------------------------- */
/* Scalars of `_sample_spikeslab` (pyblip/linear/_linear.pyx:34-53), 1:1. */
typedef struct blipgpu_spikeslab_config {
    double tau2;   int32_t update_tau2;   double tau2_a0, tau2_b0;
    double sigma2; int32_t update_sigma2; double sigma2_a0, sigma2_b0;
    double p0;     int32_t update_p0;     double min_p0, p0_a0, p0_b0;
} blipgpu_spikeslab_config;

/* Optional injected random stream (HOST arrays, chain-major).  A NULL member
 * means "use the keyed Philox stream for that slot".  Indices: c chain (local),
 * i sweep in [0, n_sweeps_total), s position in the sweep's visiting order. */
typedef struct blipgpu_streams {
    const int32_t* perm;      /* [chains][sweeps][p]  visiting order              */
    const double* u;          /* [chains][sweeps][p]  spike/slab uniform at s     */
    const double* z;          /* [chains][sweeps][p]  slab normal at s            */
    const double* invgamma;   /* [chains][sweeps]     InvGamma(n/2+sigma2_a0)     */
    const double* p0_draw;    /* [chains][sweeps][nprop] Beta variates            */
    const double* gamma_draw; /* [chains][sweeps]     Gamma(tau2_a0+k/2) variates */
    int32_t nprop;            /* 1, or 100 when min_p0 > 0                        */
} blipgpu_streams;

#define BLIPGPU_MODE_AUTO 0
#define BLIPGPU_MODE_RESIDUAL 1  /* state r = y - X beta; one X column per update  */
#define BLIPGPU_MODE_GRAM 2      /* state q = X^T r; one Gram column per change    */

#define BLIPGPU_GRAM_KERNEL_AUTO 0        /* by chains per device: specialised (<= SMs), else classic              */
#define BLIPGPU_GRAM_KERNEL_CLASSIC 1     /* one 256-thread CTA per chain, two chains per SM, work queue           */
#define BLIPGPU_GRAM_KERNEL_SPECIALIZED 2 /* one SM per chain: decision warps + helper warps in one CTA            */
#define BLIPGPU_GRAM_KERNEL_CLUSTER 3     /* two SMs per chain: a 2-CTA cluster, decisions | column stream (DSMEM) */
#define BLIPGPU_GRAM_KERNEL_BATCHED 4     /* one SM per chain: batched decisions (resolver warp + evaluators) + helper warps */
#define BLIPGPU_GRAM_KERNEL_BATCHED_CLUSTER 5 /* the batched kernel on a 2-CTA cluster per chain                   */
#define BLIPGPU_GRAM_KERNEL_STREAM 6      /* two SMs per chain: resolver warp | checked column stream (stationary regime) */
#define BLIPGPU_GRAM_KERNEL_SPECIALIZED2 7 /* SPECIALIZED with two tracked positions per thread and Gram entries requested a change ahead */
#define BLIPGPU_GRAM_KERNEL_SPECIALIZED2_CLUSTER 8 /* the same on a 2-CTA cluster: q on both SMs, a control warp does the fences */

typedef struct blipgpu_sample_options {
    int32_t mode;            /* BLIPGPU_MODE_*                                     */
    int32_t chunk_sweeps;    /* sweeps per kernel launch (0 = default)             */
    int32_t keep_latents;    /* probit: keep Z on the device                       */
    int32_t gram_refresh;    /* gram mode: rebuild q every this many sweeps (0=def) */
    int32_t gram_kernel;     /* BLIPGPU_GRAM_KERNEL_*: both kernels give bit-identical results */
    int32_t reserved[3];
} blipgpu_sample_options;

/* Replaces `apply_pool(_sample_spikeslab, constant_inputs, N=[N+burn]*chains)`
 * (pyblip/linear/linear.py:159-164, pyblip/probit/probit.py:89-94): all chains in
 * one call.  `y` is the response (linear) and ignored when probit != 0; `z` are
 * the 0/1 labels (probit; 1 => latent < 0).  Chains get global ids
 * chain_offset .. chain_offset+chains-1 (the Philox key), so a multi-GPU run
 * that shards chains produces the same draws as a single-GPU run.
 * The first `burn` sweeps are run but not recorded. */
int blipgpu_sample_spikeslab(blipgpu_ctx* ctx, blipgpu_design* design,
                             const blipgpu_spikeslab_config* cfg,
                             const double* y, const int64_t* z, int32_t probit,
                             int64_t chains, int64_t chain_offset,
                             int64_t n_keep, int64_t burn, uint64_t seed,
                             const blipgpu_streams* streams,
                             const blipgpu_sample_options* opts,
                             blipgpu_samples** out);

/* ---- block spike-and-slab sampler (bsize > 1) ---------------------------- */
/* Optional injected stream of the block sampler (HOST arrays, chain-major; NULL member =
 * keyed Philox stream).  nblock = ceil(p / bsize); a block is `bsize` consecutive indices
 * (wrapping at p) starting at a multiple of bsize. */
typedef struct blipgpu_multi_streams {
    const int32_t* blocks;    /* [chains][nblock][bsize] block table in visiting order (the
                                 reference shuffles it once, in sweep 0, _linear_multi.pyx:426) */
    const double* u;          /* [chains][sweeps][nblock]        uniform of _weighted_choice (:71)  */
    const double* z;          /* [chains][sweeps][nblock][bsize] normals of the drawn model (:768)  */
    const double* invgamma;   /* [chains][sweeps]                                                    */
    const double* p0_draw;    /* [chains][sweeps][nprop]                                             */
    const double* gamma_draw; /* [chains][sweeps]                                                    */
    int32_t nprop;
} blipgpu_multi_streams;

/* Replaces `apply_pool(_sample_spikeslab_multi, constant_inputs, N=[N+burn]*chains)`
 * (pyblip/linear/linear.py:149-164, pyblip/probit/probit.py:79-94; kernel
 * pyblip/linear/_linear_multi.pyx:271-292): all chains of this rank in one call.
 * 2 <= bsize <= 8 (larger blocks are rejected with BLIPGPU_ERR_UNSUPPORTED);
 * max_signals_per_block = 0 means no cap (:403-404).  A failed Cholesky factorisation
 * (reference: RuntimeError "dpotrf exited with INFO=...") returns BLIPGPU_ERR_NUMERIC.
 * Everything else as blipgpu_sample_spikeslab; the result is the same samples handle. */
int blipgpu_sample_spikeslab_multi(blipgpu_ctx* ctx, blipgpu_design* design,
                                   const blipgpu_spikeslab_config* cfg, int32_t bsize,
                                   int32_t max_signals_per_block,
                                   const double* y, const int64_t* z, int32_t probit,
                                   int64_t chains, int64_t chain_offset,
                                   int64_t n_keep, int64_t burn, uint64_t seed,
                                   const blipgpu_multi_streams* streams,
                                   const blipgpu_sample_options* opts,
                                   blipgpu_samples** out);

/* ---- neuronized-prior sampler ------------------------------------------ */
/* Scalars of `_nprior_sample` (pyblip/nprior/_nprior.pyx:56-73). */
typedef struct blipgpu_nprior_config {
    double tauw2, p0_init, min_p0; int32_t update_p0;
    double sigma_a0, sigma_b0;
    int32_t sigma_prior_type, joint_sample_W, group_alpha_update, reserved;
} blipgpu_nprior_config;

/* Optional injected stream for the neuronized-prior sampler (HOST arrays, chain-major;
 * NULL member = keyed Philox stream).  N = n_keep + burn. */
typedef struct blipgpu_nprior_streams {
    const int32_t* perm;       /* [chains][N][p] visiting order                             */
    const double* u;           /* [chains][N][p] mixture uniform at position s              */
    const double* wz;          /* [chains][N][p] normal of the w_j update at position s     */
    const double* alpha_init;  /* [chains][p]    N(0,1) behind alphas[0] (:98)              */
    const double* w_fresh;     /* [chains][N][p] N(0,1) behind the fresh ws rows (:100)     */
    const double* sigma_init;  /* [chains]       N(0,1) behind sigma2s[0] (:117)            */
    const double* joint_all;   /* [chains][N][p] randn(p) of _sample_w (:288)               */
    const double* joint_act;   /* [chains][N][p] randn(num_active) of _sample_w, padded     */
    const double* delta_z;     /* [chains][N]    group-move normal (:500)                   */
    const double* invgamma;    /* [chains][N]                                               */
    const double* p0_draw;     /* [chains][N]    Beta variates (:465-468)                   */
} blipgpu_nprior_streams;

typedef struct blipgpu_nprior blipgpu_nprior;   /* dense alphas / ws / betas on the device */

/* Replaces `apply_pool(_nprior_sample, ...)` (pyblip/nprior/nprior.py:126-147): all chains of
 * this rank in one call.  Needs a dense design; builds the Gram matrix if necessary. */
int blipgpu_sample_nprior(blipgpu_ctx* ctx, blipgpu_design* design, const blipgpu_nprior_config* cfg,
                          const double* y, int64_t chains, int64_t chain_offset, int64_t n_keep,
                          int64_t burn, uint64_t seed, const blipgpu_nprior_streams* streams,
                          blipgpu_nprior** out);
/* which: 0 alphas, 1 ws, 2 betas ([rows][p]); 3 sigma2s, 4 alpha0s, 5 p0s ([rows]); host out */
int blipgpu_nprior_get(blipgpu_nprior* r, int32_t which, double* out);
int blipgpu_nprior_info(blipgpu_nprior* r, int64_t* rows, int64_t* p, double* sweep_ms);
int blipgpu_nprior_bits(blipgpu_nprior* r, blipgpu_bits** out);   /* betas != 0, for counting */
int blipgpu_nprior_free(blipgpu_nprior* r);

/* ---- samples handle ---------------------------------------------------- */
typedef struct blipgpu_samples_info {
    int64_t rows;        /* chains * n_keep, chain-major like linear.py:165-168   */
    int64_t p, n, chains, n_keep;
    int64_t nnz;         /* total non-zero coefficients stored                   */
    int64_t n_events;    /* probit: total latent redraw events                   */
    int32_t has_latents, mode;
    double sweep_ms;     /* device time in sweep kernels for this call           */
    int64_t sweep_launches;
    int64_t bytes_device;
    double total_ms;     /* device time of the whole call (events on the stream) */
    int64_t n_changes;   /* coordinate visits that changed a value (all sweeps)  */
    int64_t n_windows;   /* speculation windows evaluated (all sweeps)           */
    int64_t n_sweeps_total; /* chains * (n_keep + burn)                          */
} blipgpu_samples_info;

int blipgpu_samples_info_get(blipgpu_samples* s, blipgpu_samples_info* info);
/* p0s / tau2s / sigma2s, each [rows] float64 on the host. */
int blipgpu_samples_hyper(blipgpu_samples* s, double* p0s, double* tau2s, double* sigma2s);
/* Dense betas rows [row0, row1) -> out[(row1-row0)][p] float64 host
 * (what `np.concatenate([x['betas'][burn:] ...])` returns, linear.py:165). */
int blipgpu_samples_dense(blipgpu_samples* s, int64_t row0, int64_t row1, double* out);
/* CSR export: indptr[rows+1], then indices/values[nnz] (host). */
int blipgpu_samples_csr(blipgpu_samples* s, int64_t* indptr, int32_t* indices, double* values);
/* Probit latents Z rows [row0,row1) -> out[..][n] (probit.py:99). */
int blipgpu_samples_latents(blipgpu_samples* s, int64_t row0, int64_t row1, double* out);
int blipgpu_samples_free(blipgpu_samples* s);

/* ---- candidate-group PIP counting -------------------------------------- */
/* Bit-packed (samples != 0).  Replaces `samples = samples != 0`
 * (pyblip/create_groups.py:104, 322, 435). */
int blipgpu_bits_from_samples(blipgpu_samples* s, int32_t zero_first_column, blipgpu_bits** out);
#define BLIPGPU_DTYPE_F64 0
#define BLIPGPU_DTYPE_U8 1
/* host (N x p) row-major array of float64 or uint8/bool */
int blipgpu_bits_from_dense(blipgpu_ctx* ctx, const void* samples, int32_t dtype, int64_t N,
                            int64_t p, blipgpu_bits** out);
int blipgpu_bits_shape(blipgpu_bits* b, int64_t* N, int64_t* p);
/* samples[:, cols]  (create_groups.py:117) -- compacted index space */
int blipgpu_bits_select_columns(blipgpu_bits* b, const int64_t* cols, int64_t ncols,
                                blipgpu_bits** out);
/* samples[:, 0] = 0 in place (pyblip/changepoint.py:7: time 0 is never a change point). */
int blipgpu_bits_clear_first_column(blipgpu_bits* b);
/* Unpack to a host uint8 (N x p) matrix (tests / fallbacks). */
int blipgpu_bits_to_dense(blipgpu_bits* b, uint8_t* out);

/* counts[j] = #rows with samples[:, j] != 0      (create_groups.py:105-106).
 * If out_is_device != 0, `counts` is a device pointer (e.g. a torch tensor that
 * is all-reduced with NCCL afterwards); otherwise a host pointer. */
int blipgpu_count_columns(blipgpu_bits* b, int64_t* counts, int32_t out_is_device);
/* zero_counts[m][j] = #rows whose columns j..j+m are all zero, m < max_size,
 * j < p - m; entries with j >= p - m are 0.  Layout [max_size][p].
 * (create_groups.py:321-332: cum_diffs == 0 summed over rows.) */
int blipgpu_count_sequential(blipgpu_bits* b, int32_t max_size, int64_t* zero_counts,
                             int32_t out_is_device);
/* any_counts[g] = #rows with any(samples[:, members[offsets[g]:offsets[g+1]]])
 * (create_groups.py:459, blip.py:270-275).  offsets/members are host arrays. */
int blipgpu_count_groups(blipgpu_bits* b, const int64_t* offsets, const int64_t* members,
                         int64_t n_groups, int64_t* any_counts, int32_t out_is_device);
/* pair_counts[a][b] = #rows with columns a and b both non-zero, layout [p][p], symmetric, the
 * diagonal holds the column counts.  The integer content of `np.corrcoef(precorr.T)` in
 * `_samples_dist_matrix` (create_groups.py:523-535): the correlation matrix follows from these
 * counts and N exactly.  From 512 columns up the counts are an integer GEMM on the tensor cores
 * (tcgen05.mma kind::i8 over the column-major copy of the bit matrix, any p); the scalar kernel for
 * smaller sets stops at p = 16384. */
int blipgpu_count_pairs(blipgpu_bits* b, int64_t* pair_counts, int32_t out_is_device);
/* *count = #rows in which at least one of the given groups has no non-zero column: the exact
 * family-wise error count of a selection of groups (pyblip/blip.py:270-276, evaluated up to 100
 * times by the FWER binary search).  offsets/members as in blipgpu_count_groups. */
int blipgpu_count_false_rows(blipgpu_bits* b, const int64_t* offsets, const int64_t* members,
                             int64_t n_groups, int64_t* count, int32_t out_is_device);
int blipgpu_bits_free(blipgpu_bits* b);

/* ---- continuous-space candidate regions -------------------------------- */
typedef struct blipgpu_gridcounts blipgpu_gridcounts;
/* Replaces the per-sample Python loops of `grid_peps` (pyblip/create_groups_cts.py:152-206).
 * locs: HOST [N][n_src][d] float64 in [0,1], NaN rows = no signal.  For every grid size the
 * cell containing each signal (and, if circle != 0 and d == 2, the axis-neighbour circles that
 * also contain it, :47-66) is found, de-duplicated per sample and counted over samples.
 * Optional extra centres (:165-184) with their per-grid radii.  Result: integer codes
 *   bit 63 extra-centre flag | bits 56-62 grid index | low 56 bits: (cell coordinate + 1) packed
 *   in 56/d bits per dimension (d <= 3), or the extra-centre index,
 * with the number of samples containing each code; if want_pairs != 0 additionally one
 * (code, multiplicity) pair per (sample, code) for `count_signals=True`. */
int blipgpu_grid_count(blipgpu_ctx* ctx, const double* locs, int64_t N, int64_t n_src, int32_t d,
                       const double* grid_sizes, int32_t G, int32_t circle,
                       const double* extra_centers, int64_t n_extra, const double* extra_radii,
                       const uint64_t* extra_codes /* [n_extra][G]: code counted for (centre, grid) */,
                       int32_t want_pairs, blipgpu_gridcounts** out);
int blipgpu_gridcounts_sizes(blipgpu_gridcounts* r, int64_t* n_keys, int64_t* n_pairs);
int blipgpu_gridcounts_get(blipgpu_gridcounts* r, uint64_t* codes, uint32_t* counts,
                           uint64_t* pair_codes, uint32_t* pair_mult);
int blipgpu_gridcounts_free(blipgpu_gridcounts* r);

/* Overlap (constraint) matrix of G candidate regions, the O(G^2) step of `grid_peps_to_cand_groups`
 * (pyblip/create_groups_cts.py:283-308), in the reference's float32 arithmetic.  centers: HOST
 * [d][G] float32, radii: HOST [G] float32; circle == 0: squares (all |c_a - c_b| < r_a + r_b),
 * else Euclidean distance < r_a + r_b.  out_bits: HOST [G][ceil(G/32)] uint32, bit b of word
 * [a][b/32] set iff regions a and b overlap (the diagonal is set). */
int blipgpu_overlap_bits(blipgpu_ctx* ctx, const float* centers, const float* radii, int64_t G,
                         int32_t d, int32_t circle, uint32_t* out_bits);

/* ---- hardware probes (measurement only) ----------------------------------- */
/* bench.py prints these next to its results so that a roofline fraction has a denominator measured in
 * the same process under the same clocks.
 * read_bandwidth: GB/s of 16-byte streaming loads over `bytes` for `passes` passes (after one warm-up
 *   pass): a buffer well below the 126 MB L2 measures L2 -> SM bandwidth, one far above it HBM reads.
 * fp64: TFLOP/s of kind 0 = DFMA (fp64 pipe), 1 = DMMA mma.sync.m8n8k4.f64 (the Gram build's instruction).
 * latency: cycles per dependent load of a pointer chase over `bytes` (L2-resident when small). */
int blipgpu_probe_read_bandwidth(blipgpu_ctx* ctx, int64_t bytes, int32_t passes, double* gb_per_s);
int blipgpu_probe_fp64(blipgpu_ctx* ctx, int32_t kind, double* tflops);
int blipgpu_probe_latency(blipgpu_ctx* ctx, int64_t bytes, double* cycles_per_load);

/* ---- parity hooks ---------------------------------------------------------- */
/* The device arithmetic of `sample_truncnorm` (pyblip/cython_utils/_truncnorm.pyx:42-107) on
 * RECORDED variate sequences instead of the keyed Philox stream: call i consumes uniforms
 * u[u_at[i]...] in the order of the reference's rand()/RAND_MAX calls (:50-53) and normals
 * z[z_at[i]...] in the order of its np.random.randn() calls (:63); u_used[i] / z_used[i] return
 * how many it took.  All pointers are HOST arrays.  Used by the tests to compare the device
 * with values produced by the reference's own compiled sampler (tests/golden/
 * truncnorm_values.npz, written by oracle/pin_truncnorm.py); the samplers never call it. */
int blipgpu_debug_truncnorm(blipgpu_ctx* ctx, int64_t count, const double* mean, const double* var,
                            const double* b, const int32_t* lower_interval, const double* u,
                            int64_t nu, const int64_t* u_at, const double* z, int64_t nz,
                            const int64_t* z_at, double* out, int64_t* u_used, int64_t* z_used);

#ifdef __cplusplus
}
#endif
#endif /* BLIPGPU_H */
