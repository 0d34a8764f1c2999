// gibbs.cu -- host orchestration of the multi-chain spike-and-slab sampler.
//
// Replaces the reference's chain fan-out `apply_pool(_sample_spikeslab, ...)`
// (/root/reference/pyblip/utilities.py:63-139, linear/linear.py:159-164,
// probit/probit.py:89-94): all chains of this rank run as one grid, one CTA per
// chain, in launches of `chunk_sweeps` sweeps.  Permutations for the next chunk
// are generated on a second stream while the current chunk sweeps.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>
#include "gibbs_kernels.cuh"
#include "gibbs_block.cuh"
#include "gibbs_gram_ws.cuh"
#include "gibbs_gram_batch.cuh"
#include "gibbs_gram_stream.cuh"
#include "gibbs_gram_ws2.cuh"

namespace {

// ---------------------------------------------------------------------------
// visiting orders: one block per (chain, sweep) evaluates the keyed permutation point-wise
// (common.cuh: perm_at) -- embarrassingly parallel, coalesced stores.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void perm_fill_kernel(T* perm, int chains, int S, int p, int sweep0, uint32_t k0,
                                 uint32_t k1, uint32_t chain_offset)
{
    const int64_t t = blockIdx.x;                  // (chain, sweep) pair
    const int c = (int)(t / S), s = (int)(t % S);
    const PermKey K = perm_key(k0, k1, chain_offset + (uint32_t)c, (uint32_t)(sweep0 + s), (uint32_t)p);
    T* a = perm + t * p;
    for (int pos = threadIdx.x; pos < p; pos += blockDim.x) a[pos] = (T)perm_at(K, (uint32_t)pos, (uint32_t)p);
}

// q0[j] = X_j . y   (one warp per column)
__global__ void xty_kernel(const double* __restrict__ XT, const double* __restrict__ y,
                           double* __restrict__ q0, int64_t n, int64_t p, int64_t ldx)
{
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const double* col = XT + j * ldx;
    double s = 0.0;
    for (int64_t i = threadIdx.x & 31; i < n; i += 32) s = fma(col[i], y[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) q0[j] = s;
}

// initial chain state: beta = 0 (memset), r = y or first probit latents (:117-124)
__global__ void init_state_kernel(SweepParams P, const double* __restrict__ y, int probit, int gram)
{
    const int c = blockIdx.x, tid = threadIdx.x;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    if (gram) {
        for (int k = tid; k < P.p2; k += blockDim.x) P.q[(int64_t)c * P.p2 + k] = k < P.p ? P.q0[k] : 0.0;
    } else {
        for (int k = tid; k < P.n2; k += blockDim.x) {
            double yv = 0.0, rv = 0.0;
            if (k < P.n) {
                if (probit) {
                    yv = truncnorm(key, KIND_TRUNCNORM, (uint32_t)k, 0u, 0.0, P.cfg.sigma2, 0.0, P.zlab[k]);
                    rv = yv;             // r = y - mu with mu = 0
                } else {
                    rv = y[k];
                }
            }
            P.r[(int64_t)c * P.n2 + k] = rv;
            if (probit) { P.mu[(int64_t)c * P.n2 + k] = 0.0; P.ylat[(int64_t)c * P.n2 + k] = yv; }
        }
    }
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = P.cfg.sigma2; P.hstate[GB_HS * c + 1] = P.cfg.tau2;
        P.hstate[GB_HS * c + 2] = P.cfg.p0;     P.hstate[GB_HS * c + 3] = P.yy;
        P.hstate[GB_HS * c + 4] = 0.0;          P.hstate[GB_HS * c + 5] = 0.0;
        if (probit) { P.events[2 * c] = 1u; P.events[2 * c + 1] = 1u; }
    }
}

// block sampler: block table in the keyed visiting order (the reference shuffles the rows of the
// partition once, in sweep 0: _linear_multi.pyx:395-402, 426), one thread per (chain, position)
__global__ void block_table_kernel(int32_t* blocks, int chains, int nblock, int bsize, int p, uint32_t k0,
                                   uint32_t k1, uint32_t chain_offset)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)chains * nblock) return;
    const int c = (int)(t / nblock), bi = (int)(t % nblock);
    const PermKey K = perm_key(k0, k1, chain_offset + (uint32_t)c, 0u, (uint32_t)nblock);
    const int64_t b = perm_at(K, (uint32_t)bi, (uint32_t)nblock);
    for (int k = 0; k < bsize; ++k) blocks[t * bsize + k] = (int32_t)((b * bsize + k) % p);
}

// block sampler: gb[b][a][a2] = X_ja . X_ja2 for the members of original block b (one CTA per block,
// one warp per entry); replaces the XTX[j, j2] gathers of _precompute_suff_stats (:117-121)
__global__ void block_gram_kernel(const double* __restrict__ XT, double* __restrict__ gb, int64_t n, int64_t p,
                                  int64_t ldx, int bsize)
{
    const int64_t b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int e = warp; e < bsize * bsize; e += nwarp) {
        const int64_t ja = (b * bsize + e / bsize) % p, jb = (b * bsize + e % bsize) % p;
        const double* xa = XT + ja * ldx;
        const double* xb = XT + jb * ldx;
        double s = 0.0;
        for (int64_t i = lane; i < n; i += 32) s = fma(xa[i], xb[i], s);
        s = warp_sum(s);
        if (lane == 0) gb[b * bsize * bsize + e] = s;
    }
}

// the same for the change-point design X[i][j] = 1{i >= j}: X_a . X_b = T - max(a, b)
__global__ void block_gram_changepoint_kernel(double* __restrict__ gb, int64_t T, int64_t nblock, int bsize)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblock * bsize * bsize) return;
    const int64_t b = t / (bsize * bsize);
    const int e = (int)(t % (bsize * bsize));
    const int64_t ja = (b * bsize + e / bsize) % T, jb = (b * bsize + e % bsize) % T;
    gb[t] = (double)(T - (ja > jb ? ja : jb));
}

template <class T>
int upload_chunk(T* dst, const T* src_host, int64_t chains, int64_t total_per_chain,
                 int64_t off_per_chain, int64_t count_per_chain, cudaStream_t st)
{   // host [chains][total] -> device [chains][count], starting at `off` inside every chain
    CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(T) * count_per_chain, src_host + off_per_chain,
                               sizeof(T) * total_per_chain, sizeof(T) * count_per_chain,
                               (size_t)chains, cudaMemcpyHostToDevice, st));
    return BLIP_OK;
}

} // namespace

// helpers shared with the other samplers (declared in internal.cuh)
int blip_fill_perms(cudaStream_t st, void* buf, int perm16, int64_t chains, int S, int64_t p,
                    int64_t sweep0, uint32_t k0, uint32_t k1, uint32_t chain_offset)
{
    const int64_t nt = chains * S;
    if (nt <= 0) return BLIP_OK;
    if (perm16)
        perm_fill_kernel<uint16_t><<<(unsigned)nt, 256, 0, st>>>((uint16_t*)buf, (int)chains, S, (int)p, (int)sweep0, k0, k1, chain_offset);
    else
        perm_fill_kernel<int32_t><<<(unsigned)nt, 256, 0, st>>>((int32_t*)buf, (int)chains, S, (int)p, (int)sweep0, k0, k1, chain_offset);
    CUDA_TRY(cudaGetLastError());
    return BLIP_OK;
}

int blip_xty(cudaStream_t st, const blipgpu_design* d, const double* d_y, double* d_q0)
{
    xty_kernel<<<(unsigned)((d->p + 7) / 8), 256, 0, st>>>(d->XT, d_y, d_q0, d->n, d->p, d->ldx);
    CUDA_TRY(cudaGetLastError());
    return BLIP_OK;
}

// Shared implementation of the single-site (bsize == 1) and block (bsize > 1) samplers.  For the
// block sampler `streams` carries the hyper-parameter variates and u / z with nblock resp.
// nblock*bsize entries per sweep, and `blocks_host` the optional injected block table.
static int sample_impl(blipgpu_ctx* ctx, blipgpu_design* design,
                       const blipgpu_spikeslab_config* cfg, const double* y,
                       const int64_t* z, int32_t probit, int64_t chains,
                       int64_t chain_offset, int64_t n_keep, int64_t burn,
                       uint64_t seed, const blipgpu_streams* streams,
                       const blipgpu_sample_options* opts, blipgpu_samples** out,
                       int bsize, int max_signals, const int32_t* blocks_host)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!design || !cfg || !out || chains <= 0 || n_keep <= 0 || burn < 0) {
        blip_set_error("sample_spikeslab: bad arguments"); return BLIP_ERR_ARG;
    }
    if (design->ctx != ctx) { blip_set_error("sample_spikeslab: design belongs to another context"); return BLIP_ERR_ARG; }
    if (probit ? !z : !y) { blip_set_error("sample_spikeslab: %s is NULL", probit ? "z" : "y"); return BLIP_ERR_ARG; }
    const int64_t n = design->n, p = design->p;
    const int64_t n_total = n_keep + burn;
    if (n_total >= (int64_t)1 << 31 || chains * n_keep >= (int64_t)1 << 40) {
        blip_set_error("sample_spikeslab: too many sweeps"); return BLIP_ERR_ARG;
    }
    const bool block = bsize > 1;
    const int64_t nblock = block ? (p + bsize - 1) / bsize : 0;     // _linear_multi.pyx:298 (C integer division)
    std::vector<uint32_t> model_masks;
    if (block) {
        if (bsize > BK_MAXB || bsize > p) {
            blip_set_error("sample_spikeslab_multi: bsize %d is not supported (2 <= bsize <= min(%d, p))", bsize, BK_MAXB);
            return BLIP_ERR_UNSUPPORTED;
        }
        if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT && probit) {
            blip_set_error("sample_spikeslab_multi: the structured change-point design is linear only"); return BLIP_ERR_UNSUPPORTED;
        }
        const int cap = max_signals <= 0 ? bsize : std::min(max_signals, bsize);   // :403-404
        for (int m = 1; m <= cap; ++m) {                            // itertools.combinations order (:405-409)
            std::vector<int> idx((size_t)m);
            for (int t = 0; t < m; ++t) idx[(size_t)t] = t;
            for (;;) {
                uint32_t mask = 0;
                for (int t = 0; t < m; ++t) mask |= 1u << idx[(size_t)t];
                model_masks.push_back(mask);
                int t = m - 1;
                while (t >= 0 && idx[(size_t)t] == bsize - m + t) --t;
                if (t < 0) break;
                idx[(size_t)t]++;
                for (int q2 = t + 1; q2 < m; ++q2) idx[(size_t)q2] = idx[(size_t)q2 - 1] + 1;
            }
        }
        if (blocks_host) {
            for (int64_t i = 0; i < chains * nblock; ++i) {
                const int32_t j0 = blocks_host[i * bsize];
                bool ok = j0 >= 0 && j0 < p && j0 % bsize == 0;
                for (int k = 1; ok && k < bsize; ++k) ok = blocks_host[i * bsize + k] == (int32_t)((j0 + k) % p);
                if (!ok) { blip_set_error("sample_spikeslab_multi: injected block %lld is not bsize consecutive indices starting at a multiple of bsize", (long long)i); return BLIP_ERR_ARG; }
            }
        }
    }
    const int nmodel = (int)model_masks.size() + 1;
    int mode = opts ? opts->mode : BLIPGPU_MODE_AUTO;
    if (mode == BLIPGPU_MODE_AUTO) {
        if (probit) mode = BLIPGPU_MODE_RESIDUAL;
        else if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT) mode = BLIPGPU_MODE_GRAM;
        else {
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            const double gram_bytes = 8.0 * (double)p * (double)round_up(p, 4);
            mode = (design->G || gram_bytes < 0.5 * (double)free_b) ? BLIPGPU_MODE_GRAM : BLIPGPU_MODE_RESIDUAL;
        }
    }
    if (block && mode == BLIPGPU_MODE_GRAM && (!opts || opts->mode == BLIPGPU_MODE_AUTO)) {
        // the block kernel keeps q in shared memory; fall back to the residual form if it cannot
        const size_t need = sizeof(double) * ((size_t)round_up(p, 2) + (size_t)GB_NWARP * std::max(32, (nmodel + 1) & ~1)) +
                            sizeof(uint32_t) * (size_t)((p + 31) / 32);
        if (need > ctx->smem_optin) mode = BLIPGPU_MODE_RESIDUAL;
    }
    if (probit && mode == BLIPGPU_MODE_GRAM) {
        blip_set_error("sample_spikeslab: probit needs residual mode (every change redraws all latents)");
        return BLIP_ERR_ARG;
    }
    if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT && mode != BLIPGPU_MODE_GRAM) {
        blip_set_error("sample_spikeslab: the structured change-point design only supports gram mode");
        return BLIP_ERR_ARG;
    }
    const bool gram = mode == BLIPGPU_MODE_GRAM;
    if (gram && design->kind == BLIPGPU_DESIGN_DENSE) BLIP_TRY(blipgpu_design_build_gram(design));

    const int64_t words = (p + 31) / 32, n2 = round_up(n, 2), p2 = round_up(p, 2);

    // ---- shared memory budget -> kernel variant
    size_t smem = 0;
    bool qsmem = true;
    if (block) {
        const size_t state = gram ? (size_t)p2 : (size_t)(probit ? 3 : 1) * n2;
        smem = sizeof(double) * (state + (size_t)GB_NWARP * std::max(32, (nmodel + 1) & ~1)) + sizeof(uint32_t) * words;
    } else if (gram) {
        smem = sizeof(double) * p2 + sizeof(uint32_t) * words;
        if (smem > ctx->smem_optin) { qsmem = false; smem = sizeof(uint32_t) * words; }
    } else {
        smem = sizeof(double) * ((probit ? 3 : 1) * n2 + GB_NWARP * 32 + 40) + sizeof(uint32_t) * words;
    }
    if (smem > ctx->smem_optin) {
        blip_set_error("sample_spikeslab: problem too large for %s mode (needs %zu bytes of shared "
                       "memory per chain, device allows %zu)", gram ? "gram" : "residual", smem, ctx->smem_optin);
        return BLIP_ERR_UNSUPPORTED;
    }

    // ---- few chains: role-specialised kernel, one SM per chain or a 2-CTA cluster per chain (gibbs_gram_ws.cuh)
    const int want_kernel = opts ? opts->gram_kernel : BLIPGPU_GRAM_KERNEL_AUTO;
    const size_t smem_ws = 2 * sizeof(double) * p2 + sizeof(uint32_t) * words;      // q twice + active-set mask
    const bool ws_fits = gram && !block && smem_ws + 8192 <= ctx->smem_optin;
    const bool ws_possible = ws_fits && chains <= ctx->sm_count;
    const bool cl_possible = ws_fits && 2 * chains <= ctx->sm_count;
    if ((want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED && !ws_possible) || (want_kernel == BLIPGPU_GRAM_KERNEL_CLUSTER && !cl_possible)) {
        blip_set_error("sample_spikeslab: the role-specialised gram kernels need gram mode, bsize = 1, 16 p bytes of shared "
                       "memory and at most %d chains per device (%d for the cluster placement)", ctx->sm_count, ctx->sm_count / 2);
        return BLIP_ERR_UNSUPPORTED;
    }
    // AUTO never picks the cluster placement: measured SLOWER than one SM per chain (6.8 vs 5.7 ms per
    // 32 sweeps of 64 chains at C2; classic 6.4): a distributed-shared-memory read of q is served by the
    // helper SM's load/store pipeline, behind the helpers' own column stream.
    // BLIPGPU_WS_CLUSTER=1 or gram_kernel = CLUSTER selects it.
    const bool no_ws = getenv("BLIPGPU_NO_WS") != nullptr, env_cl = getenv("BLIPGPU_WS_CLUSTER") != nullptr;
    // batched decisions (gibbs_gram_batch.cuh): the same two placements, plus the visiting order in shared memory
    const size_t smem_wb = smem_ws + (size_t)round_up(sizeof(uint16_t) * p, 16);
    const bool wb_fits = gram && !block && p <= 65535 && smem_wb + 12288 <= ctx->smem_optin;
    const bool wb_possible = wb_fits && chains <= ctx->sm_count, wbc_possible = wb_fits && 2 * chains <= ctx->sm_count;
    if ((want_kernel == BLIPGPU_GRAM_KERNEL_BATCHED && !wb_possible) || (want_kernel == BLIPGPU_GRAM_KERNEL_BATCHED_CLUSTER && !wbc_possible)) {
        blip_set_error("sample_spikeslab: the batched gram kernels need gram mode, bsize = 1, 18 p bytes of shared "
                       "memory and at most %d chains per device (%d for the cluster placement)", ctx->sm_count, ctx->sm_count / 2);
        return BLIP_ERR_UNSUPPORTED;
    }
    // checked column stream (gibbs_gram_stream.cuh): q twice, the tracked Gram sub-matrix, the resolver's snapshots
    const size_t smem_st = st_smem_bytes(p2, words);
    const bool st_fits = gram && !block && smem_st + sizeof(StCtl) + sizeof(StPass) + 1024 <= ctx->smem_optin;
    const bool st_possible = st_fits && chains <= ctx->sm_count;
    if (want_kernel == BLIPGPU_GRAM_KERNEL_STREAM && !st_possible) {
        blip_set_error("sample_spikeslab: the column-stream gram kernel needs gram mode, bsize = 1, %zu bytes of shared memory "
                       "per chain (device allows %zu) and at most %d chains per device", smem_st + sizeof(StCtl) + sizeof(StPass), ctx->smem_optin, ctx->sm_count);
        return BLIP_ERR_UNSUPPORTED;
    }
    if (want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED2_CLUSTER && !cl_possible) {
        blip_set_error("sample_spikeslab: the 2-CTA cluster placement needs gram mode, bsize = 1, 16 p bytes of shared "
                       "memory and at most %d chains per device", ctx->sm_count / 2);
        return BLIP_ERR_UNSUPPORTED;
    }
    if (want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED2 && !ws_possible) {
        blip_set_error("sample_spikeslab: the role-specialised gram kernels need gram mode, bsize = 1, 16 p bytes of shared "
                       "memory and at most %d chains per device", ctx->sm_count);
        return BLIP_ERR_UNSUPPORTED;
    }
    bool use_cl = false, use_ws = false, use_wb = false, use_wbc = false, use_ws2 = false, use_ws2c = false;
    if (want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED2) { use_ws = true; use_ws2 = true; }
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED2_CLUSTER) use_ws2c = true;
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_CLUSTER) use_cl = true;
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_SPECIALIZED) use_ws = true;
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_BATCHED) use_wb = true;
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_BATCHED_CLUSTER) use_wbc = true;
    else if (want_kernel == BLIPGPU_GRAM_KERNEL_AUTO && !no_ws) {
        const char* env_wb = getenv("BLIPGPU_WB");
        if (env_wb && env_wb[0] == '2' && wbc_possible) use_wbc = true;
        else if (env_wb && env_wb[0] == '1' && wb_possible) use_wb = true;
        else if (cl_possible && env_cl) use_cl = true;
        else if (cl_possible && getenv("BLIPGPU_WS2C") != nullptr) use_ws2c = true;
        else if (ws_possible) { use_ws = true; use_ws2 = getenv("BLIPGPU_WS2") != nullptr; }
    }
    // The column-stream kernel is a stationary-regime kernel: forced by gram_kernel = STREAM, or taking over from the
    // role-specialised kernel launch by launch, once a launch has seen few enough changes per sweep for the resolver
    // to track them (chain state crosses launches in global memory; all kernels agree bit for bit)
    const bool st_forced = want_kernel == BLIPGPU_GRAM_KERNEL_STREAM;
    // (AUTO hands over only on request, BLIPGPU_STREAM=1: at C2 the column-stream kernel measures 12.4 ms per 32 sweeps of
    // 64 chains against 5.7 ms for the role-specialised kernel)
    const bool st_auto = use_ws && st_possible && 2 * chains <= ctx->sm_count && want_kernel == BLIPGPU_GRAM_KERNEL_AUTO && getenv("BLIPGPU_STREAM") != nullptr;
    const bool use_role = use_cl || use_ws || use_wb || use_wbc || use_ws2c || st_forced;   // double-buffered scratch, no work queue

    const auto t_call0 = std::chrono::steady_clock::now();
    // ---- chunking
    // Sweeps per launch.  The gram-form kernels compute the visiting order themselves (no permutation buffers) and
    // lose the tail of every launch (the last work items of a persistent grid, the last sweeps of the slowest chain):
    // 128 sweeps per launch measured 2.6 % faster than 32 at C2 (12.54 vs 12.87 ms per 32 sweeps of 512 chains).
    // The other kernels read pre-generated permutations and stay at 32, capped by the size of those buffers.
    const bool inj_perm = streams && streams->perm;
    const bool gram_default_chunk = mode == BLIPGPU_MODE_GRAM && !block && !inj_perm && !(streams && (streams->u || streams->z));
    int S = opts && opts->chunk_sweeps > 0 ? opts->chunk_sweeps : (gram_default_chunk ? 128 : 32);
    {
        const double perm_bytes_per_sweep = 4.0 * (double)chains * (double)p;
        const int cap = (int)std::max(1.0, 1.0e9 / perm_bytes_per_sweep);
        if (!(opts && opts->chunk_sweeps > 0) && !gram_default_chunk) S = std::min(S, cap);
        S = (int)std::min<int64_t>(S, n_total);
    }
    // Permutations are generated for PS sweeps at a time (a multiple of the launch chunk S):
    // 16-bit entries when p allows; <= 1.5 GB per buffer, two buffers.
    const bool perm16 = !inj_perm && p <= 65535;
    const size_t perm_elt = perm16 ? sizeof(uint16_t) : sizeof(int32_t);
    int64_t PS = S;
    if (!inj_perm) {
        const double per_sweep = (double)perm_elt * (double)chains * (double)p;
        int64_t mult = std::max<int64_t>(1, (int64_t)(1.5e9 / per_sweep) / S);
        mult = std::min<int64_t>(mult, std::max<int64_t>(1, 256 / S));
        PS = std::min<int64_t>(S * mult, round_up(n_total, S));
    }
    const bool inj_u = streams && streams->u, inj_z = streams && streams->z;
    const int64_t u_len = block ? nblock : p, z_len = block ? nblock * bsize : p;   // injected variates per sweep
    const bool inj_ig = streams && streams->invgamma, inj_p0 = streams && streams->p0_draw;
    const bool inj_g = streams && streams->gamma_draw;
    const int nprop = streams ? std::max(1, (int)streams->nprop) : 1;
    if (inj_p0 && cfg->min_p0 != 0 && nprop < 100) {
        blip_set_error("sample_spikeslab: min_p0 > 0 needs 100 injected Beta proposals per sweep");
        return BLIP_ERR_ARG;
    }

    // ---- allocate
    DevPoolScope pool_scope(ctx->stream);           // every DevBuf of this call: stream-ordered pool, on the sweep stream
    DevBuf d_beta, d_r, d_mu, d_ylat, d_q, d_h, d_mask, d_ev, d_q0, d_y, d_z, d_ovf, d_scrL, d_scrZ, d_scrP, d_queue;
    DevBuf d_perm[2], d_u, d_zn, d_ig, d_p0, d_g, d_scrI;
    DevBuf snap_beta, snap_r, snap_mu, snap_ylat, snap_q, snap_h, snap_mask, snap_ev;
    BLIP_TRY(d_beta.alloc(sizeof(double) * chains * p));
    BLIP_TRY(d_h.alloc(sizeof(double) * chains * GB_HS));
    BLIP_TRY(d_mask.alloc(sizeof(uint32_t) * chains * words));
    BLIP_TRY(d_ovf.alloc(sizeof(int)));
    BLIP_TRY(snap_beta.alloc(d_beta.bytes)); BLIP_TRY(snap_h.alloc(d_h.bytes)); BLIP_TRY(snap_mask.alloc(d_mask.bytes));
    if (gram) {
        BLIP_TRY(d_q.alloc(sizeof(double) * chains * p2)); BLIP_TRY(snap_q.alloc(d_q.bytes));
        BLIP_TRY(d_q0.alloc(sizeof(double) * p));
        BLIP_TRY(d_queue.alloc(sizeof(int) * (chains + 1)));
        BLIP_TRY(d_scrL.alloc(sizeof(float) * chains * p * (use_role ? 2 : 1))); BLIP_TRY(d_scrZ.alloc(sizeof(double2) * chains * p));
        if (st_forced || st_auto) BLIP_TRY(d_scrI.alloc(sizeof(int) * 2 * chains * p2));
    } else {
        BLIP_TRY(d_r.alloc(sizeof(double) * chains * n2)); BLIP_TRY(snap_r.alloc(d_r.bytes));
        if (probit) {
            BLIP_TRY(d_mu.alloc(d_r.bytes)); BLIP_TRY(d_ylat.alloc(d_r.bytes));
            BLIP_TRY(snap_mu.alloc(d_r.bytes)); BLIP_TRY(snap_ylat.alloc(d_r.bytes));
            BLIP_TRY(d_ev.alloc(sizeof(uint32_t) * chains * 2)); BLIP_TRY(snap_ev.alloc(d_ev.bytes));
            BLIP_TRY(d_z.alloc(sizeof(int32_t) * n));
        }
    }
    if (!probit) BLIP_TRY(d_y.alloc(sizeof(double) * n));
    // gram mode evaluates the keyed visiting order inside the sweep kernel (chain-private scratch);
    // the other kernels read permutation buffers filled ahead of them on the second stream
    const bool perm_in_kernel = gram && !inj_perm && !block;
    const bool gen_perms = !inj_perm && !perm_in_kernel && !block;
    if (perm_in_kernel) BLIP_TRY(d_scrP.alloc(perm_elt * chains * p * (use_role ? 2 : 1)));
    else if (!block) {
        BLIP_TRY(d_perm[0].alloc(perm_elt * chains * PS * p));
        if (!inj_perm && n_total > S) BLIP_TRY(d_perm[1].alloc(perm_elt * chains * PS * p));
    }
    if (inj_u) BLIP_TRY(d_u.alloc(sizeof(double) * chains * S * u_len));
    if (inj_z) BLIP_TRY(d_zn.alloc(sizeof(double) * chains * S * z_len));
    if (inj_ig) BLIP_TRY(d_ig.alloc(sizeof(double) * chains * S));
    if (inj_p0) BLIP_TRY(d_p0.alloc(sizeof(double) * chains * S * nprop));
    if (inj_g) BLIP_TRY(d_g.alloc(sizeof(double) * chains * S));
    {
        // pool allocations are ordered on the sweep stream; the second stream (permutation generation) joins there
        cudaEvent_t ev_alloc;
        CUDA_TRY(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
        cudaError_t ea = cudaEventRecord(ev_alloc, ctx->stream);
        if (ea == cudaSuccess) ea = cudaStreamWaitEvent(ctx->stream2, ev_alloc, 0);
        cudaEventDestroy(ev_alloc);
        CUDA_TRY(ea);
    }

    blipgpu_samples* smp = new (std::nothrow) blipgpu_samples();
    if (!smp) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    smp->ctx = ctx; smp->rows = chains * n_keep; smp->rows_pad = round_up(smp->rows, 32);
    smp->p = p; smp->n = n; smp->chains = chains; smp->n_keep = n_keep; smp->words = words;
    smp->bits = nullptr; smp->row_off = nullptr; smp->row_nnz = nullptr; smp->values = nullptr;
    smp->arena_used = nullptr; smp->hyper = nullptr; smp->latents = nullptr;
    smp->n_events = 0; smp->mode = mode; smp->sweep_ms = 0.0; smp->sweep_launches = 0;
    smp->total_ms = 0.0; smp->n_changes = 0; smp->n_windows = 0; smp->n_sweeps_total = chains * n_total;
    cudaEvent_t ev_call0, ev_call1;
    CUDA_TRY(cudaEventCreate(&ev_call0)); CUDA_TRY(cudaEventCreate(&ev_call1));
    struct CallEv { cudaEvent_t a, b; ~CallEv() { cudaEventDestroy(a); cudaEventDestroy(b); } } callev{ ev_call0, ev_call1 };
    CUDA_TRY(cudaEventRecord(ev_call0, ctx->stream));
    struct Guard { blipgpu_samples* s; ~Guard() { if (s) blipgpu_samples_free(s); } } guard{ smp };
    const bool keep_lat = probit && (!opts || opts->keep_latents != 0);
    const auto t_alloc0 = std::chrono::steady_clock::now();
    {
        cudaError_t e;
        cudaStream_t sa = ctx->stream;               // stream-ordered pool allocations (freed in blipgpu_samples_free)
        if ((e = blip_store_alloc((void**)&smp->bits, sizeof(uint32_t) * words * smp->rows_pad, sa)) != cudaSuccess ||
            (e = blip_store_alloc((void**)&smp->row_off, sizeof(int64_t) * smp->rows, sa)) != cudaSuccess ||
            (e = blip_store_alloc((void**)&smp->row_nnz, sizeof(int32_t) * smp->rows, sa)) != cudaSuccess ||
            (e = blip_store_alloc((void**)&smp->hyper, sizeof(double) * 3 * smp->rows, sa)) != cudaSuccess ||
            (e = blip_store_alloc((void**)&smp->arena_used, sizeof(unsigned long long), sa)) != cudaSuccess ||
            (keep_lat && (e = blip_store_alloc((void**)&smp->latents, sizeof(double) * smp->rows * n, sa)) != cudaSuccess)) {
            blip_set_error("sample_spikeslab: cannot allocate the sample store: %s", cudaGetErrorString(e));
            return BLIP_ERR_NOMEM;
        }
    }
    cudaStream_t st = ctx->stream, st2 = ctx->stream2;
    CUDA_TRY(cudaMemsetAsync(smp->bits, 0, sizeof(uint32_t) * words * smp->rows_pad, st));
    CUDA_TRY(cudaMemsetAsync(smp->arena_used, 0, sizeof(unsigned long long), st));
    if (keep_lat) CUDA_TRY(cudaMemsetAsync(smp->latents, 0, sizeof(double) * smp->rows * n, st));
    smp->arena_cap = std::max<int64_t>(chains * (int64_t)std::min<int64_t>(S, n_keep) * std::max<int64_t>(64, p / 16), 1 << 16);
    // a repeated job (same chains, sweeps, p on this context) starts with what the previous one ended up needing (+15 %):
    // no mid-run growth, and every pool request has the size of a block the previous call returned
    const bool hint_hit = ctx->arena_hint > 0 && ctx->arena_hint_key[0] == chains && ctx->arena_hint_key[1] == n_keep && ctx->arena_hint_key[2] == p;
    if (hint_hit) smp->arena_cap = std::max<int64_t>(smp->arena_cap, ctx->arena_hint);
    CUDA_TRY(blip_store_alloc((void**)&smp->values, sizeof(double) * smp->arena_cap, st));
    if (getenv("BLIPGPU_LAUNCH_LOG"))
        fprintf(stderr, "[blipgpu] sample store allocated in %.2f ms (host clock)\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_alloc0).count());

    // ---- parameters
    SweepParams P;
    memset(&P, 0, sizeof(P));
    P.XT = design->XT; P.ldx = design->ldx; P.Xl2 = design->Xl2; P.G = design->G; P.ldg = design->ldg;
    P.n = (int)n; P.p = (int)p; P.words = (int)words; P.n2 = (int)n2; P.p2 = (int)p2;
    P.cfg = *cfg;
    P.beta = d_beta.as<double>(); P.r = d_r.as<double>(); P.mu = d_mu.as<double>();
    P.ylat = d_ylat.as<double>(); P.q = d_q.as<double>(); P.hstate = d_h.as<double>();
    P.mask = d_mask.as<uint32_t>(); P.events = d_ev.as<uint32_t>();
    P.q0 = d_q0.as<double>(); P.zlab = d_z.as<int32_t>();
    P.scrL = d_scrL.as<float>(); P.scrZ = d_scrZ.as<double2>();
    P.scrP = d_scrP.ptr; P.scr_stride = chains * p; P.scrI = d_scrI.as<int>();
    P.queue = d_queue.as<int>(); P.progress = d_queue.as<int>() + 1; P.n_chains = (int)chains;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); P.chain_offset = (uint32_t)chain_offset;
    P.nprop = nprop; P.burn = (int)burn; P.n_keep = (int)n_keep;
    P.bits = smp->bits; P.rows = smp->rows; P.rows_pad = smp->rows_pad;
    P.row_off = smp->row_off; P.row_nnz = smp->row_nnz; P.values = smp->values;
    P.arena_cap = smp->arena_cap; P.arena_used = smp->arena_used; P.overflow = d_ovf.as<int>();
    P.hyper_out = smp->hyper; P.latents = smp->latents;
    P.gram_refresh = opts && opts->gram_refresh > 0 ? opts->gram_refresh : (opts && opts->gram_refresh < 0 ? 0 : 128);

    // ---- block sampler tables
    DevBuf d_blocks, d_gb, d_masks, d_numerr;
    BLIP_TRY(d_numerr.alloc(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(d_numerr.ptr, 0, sizeof(int), ctx->stream));
    P.numeric_err = d_numerr.as<int>();
    if (block) {
        BLIP_TRY(d_blocks.alloc(sizeof(int32_t) * chains * nblock * bsize));
        BLIP_TRY(d_gb.alloc(sizeof(double) * nblock * bsize * bsize));
        BLIP_TRY(d_masks.alloc(sizeof(uint32_t) * model_masks.size()));
        CUDA_TRY(cudaMemcpyAsync(d_masks.ptr, model_masks.data(), d_masks.bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (blocks_host) CUDA_TRY(cudaMemcpyAsync(d_blocks.ptr, blocks_host, d_blocks.bytes, cudaMemcpyHostToDevice, ctx->stream));
        else block_table_kernel<<<(unsigned)((chains * nblock + 255) / 256), 256, 0, ctx->stream>>>(
                d_blocks.as<int32_t>(), (int)chains, (int)nblock, bsize, (int)p, P.k0, P.k1, P.chain_offset);
        if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT)
            block_gram_changepoint_kernel<<<(unsigned)((nblock * bsize * bsize + 255) / 256), 256, 0, ctx->stream>>>(
                d_gb.as<double>(), p, nblock, bsize);
        else
            block_gram_kernel<<<(unsigned)nblock, 256, 0, ctx->stream>>>(design->XT, d_gb.as<double>(), n, p, design->ldx, bsize);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));       // model_masks / blocks_host are host temporaries
        P.blocks = d_blocks.as<int32_t>(); P.gb = d_gb.as<double>(); P.model_masks = d_masks.as<uint32_t>();
        P.bsize = bsize; P.nblock = (int)nblock; P.nmodel = nmodel;
        P.npad = 1; while (P.npad < nmodel) P.npad <<= 1;
        // blocks a warp scores at once: one lane per model, limited by the per-warp scratch
        // (64 doubles of Gram blocks, 16 of X_A^T r)
        P.blocks_per_warp = P.npad <= 32 ? std::max(1, std::min({ 32 / P.npad, 64 / (bsize * bsize), 16 / bsize })) : 1;
        ctx->kernel_launches += blocks_host ? 1 : 2;
    }

    // ---- data upload and initial state
    double yy = 0.0;
    if (!probit) {
        for (int64_t i = 0; i < n; ++i) yy += y[i] * y[i];
        CUDA_TRY(cudaMemcpyAsync(d_y.ptr, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    } else {
        std::vector<int32_t> z32((size_t)n);
        for (int64_t i = 0; i < n; ++i) z32[(size_t)i] = (int32_t)z[i];
        CUDA_TRY(cudaMemcpyAsync(d_z.ptr, z32.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    P.yy = yy;
    if (gram) {
        if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT) {
            // X_j . y = sum_{i >= j} y_i  (suffix sums)
            std::vector<double> q0h((size_t)p);
            double acc = 0.0;
            for (int64_t j = p - 1; j >= 0; --j) { acc += y[j]; q0h[(size_t)j] = acc; }
            CUDA_TRY(cudaMemcpyAsync(d_q0.ptr, q0h.data(), sizeof(double) * p, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        } else {
            xty_kernel<<<(unsigned)((p + 7) / 8), 256, 0, st>>>(design->XT, d_y.as<double>(), d_q0.as<double>(), n, p, design->ldx);
        }
    }
    CUDA_TRY(cudaMemsetAsync(d_beta.ptr, 0, d_beta.bytes, st));
    CUDA_TRY(cudaMemsetAsync(d_mask.ptr, 0, d_mask.bytes, st));
    init_state_kernel<<<(unsigned)chains, 256, 0, st>>>(P, d_y.as<double>(), probit, gram ? 1 : 0);
    CUDA_TRY(cudaGetLastError());

    // ---- kernel selection
    void (*kern)(SweepParams) = nullptr;
    if (block) kern = gram ? (design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_block_kernel<true, false, BLIPGPU_DESIGN_CHANGEPOINT>
                                                                        : gibbs_block_kernel<true, false>)
                           : (probit ? gibbs_block_kernel<false, true> : gibbs_block_kernel<false, false>);
    else if (!gram) kern = probit ? gibbs_residual_kernel<true> : gibbs_residual_kernel<false>;
    else if (design->kind == BLIPGPU_DESIGN_CHANGEPOINT)
        kern = qsmem ? gibbs_gram_kernel<BLIPGPU_DESIGN_CHANGEPOINT, true> : gibbs_gram_kernel<BLIPGPU_DESIGN_CHANGEPOINT, false>;
    else
        kern = qsmem ? gibbs_gram_kernel<BLIPGPU_DESIGN_DENSE, true> : gibbs_gram_kernel<BLIPGPU_DESIGN_DENSE, false>;
    if (use_ws2 || use_ws2c) {
        kern = design->kind == BLIPGPU_DESIGN_CHANGEPOINT
                   ? (use_ws2c ? gibbs_gram_ws2_kernel<BLIPGPU_DESIGN_CHANGEPOINT, true> : gibbs_gram_ws2_kernel<BLIPGPU_DESIGN_CHANGEPOINT, false>)
                   : (use_ws2c ? gibbs_gram_ws2_kernel<BLIPGPU_DESIGN_DENSE, true> : gibbs_gram_ws2_kernel<BLIPGPU_DESIGN_DENSE, false>);
        smem = smem_ws;
    } else if (use_ws) {
        kern = design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_gram_ws_kernel<BLIPGPU_DESIGN_CHANGEPOINT, false>
                                                          : gibbs_gram_ws_kernel<BLIPGPU_DESIGN_DENSE, false>;
        smem = smem_ws;
    } else if (use_cl) {
        kern = design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_gram_ws_kernel<BLIPGPU_DESIGN_CHANGEPOINT, true>
                                                          : gibbs_gram_ws_kernel<BLIPGPU_DESIGN_DENSE, true>;
        smem = smem_ws;
    } else if (use_wb) {
        kern = design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_gram_batch_kernel<BLIPGPU_DESIGN_CHANGEPOINT, false>
                                                          : gibbs_gram_batch_kernel<BLIPGPU_DESIGN_DENSE, false>;
        smem = smem_wb;
    } else if (use_wbc) {
        kern = design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_gram_batch_kernel<BLIPGPU_DESIGN_CHANGEPOINT, true>
                                                          : gibbs_gram_batch_kernel<BLIPGPU_DESIGN_DENSE, true>;
        smem = smem_wb;
    }
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void (*kern_st)(SweepParams) = design->kind == BLIPGPU_DESIGN_CHANGEPOINT ? gibbs_gram_stream_kernel<BLIPGPU_DESIGN_CHANGEPOINT>
                                                                              : gibbs_gram_stream_kernel<BLIPGPU_DESIGN_DENSE>;
    if (st_forced || st_auto) CUDA_TRY(cudaFuncSetAttribute(kern_st, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_st));
    // AUTO: changes per chain-sweep (mean, max over chains) of the previous launch decide which kernel runs the next one
    bool st_now = st_forced;
    std::vector<double> hs_prev, hs_now;
    const double st_mean_max = getenv("BLIPGPU_STREAM_MEAN") ? atof(getenv("BLIPGPU_STREAM_MEAN")) : 90.0;
    // gram mode: persistent CTAs (as many as are resident at once) draw (chain, 8-sweep) items
    // from a queue; with q in global memory (huge p) a chain stays on one CTA for the launch.
    int resident = 0;
    if (gram && !block && !use_role) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GB_BLOCK, smem));
        resident = std::max(1, per_sm) * ctx->sm_count;
    }
    const int sub_sweeps_opt = qsmem ? 8 : 0;
    // q in global memory (p too large for shared memory): every flush reads and writes a chain's q, so
    // q is pinned in L2 as far as the persisting carve-out goes and the rest of the traffic (Gram
    // columns) is left to stream past it
    struct L2Guard {
        cudaStream_t st; bool on;
        ~L2Guard() {
            if (!on) return;
            cudaStreamAttrValue a; memset(&a, 0, sizeof(a));
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
            cudaCtxResetPersistingL2Cache();
        }
    } l2g{ st, false };
    // residual form: all chains stream the same X every sweep, so X is what is pinned
    const void* l2_ptr = nullptr; size_t l2_bytes = 0;
    if (gram && !block && !qsmem) { l2_ptr = d_q.ptr; l2_bytes = d_q.bytes; }
    else if (!gram && design->XT && getenv("BLIPGPU_NO_L2_PIN") == nullptr) { l2_ptr = design->XT; l2_bytes = sizeof(double) * (size_t)design->ldx * (size_t)p; }
    if (l2_ptr) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        const size_t window = std::min<size_t>(l2_bytes, (size_t)std::max(0, max_window));
        const size_t persist = std::min<size_t>(window, (size_t)std::max(0, max_persist));
        if (persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
            cudaStreamAttrValue a; memset(&a, 0, sizeof(a));
            a.accessPolicyWindow.base_ptr = const_cast<void*>(l2_ptr);
            a.accessPolicyWindow.num_bytes = window;
            a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)window);
            a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess) l2g.on = true;
        }
        cudaGetLastError();
    }

    cudaEvent_t ev_perm[2], ev_done[2];
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&ev_perm[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    struct EvGuard { cudaEvent_t* a; cudaEvent_t* b; ~EvGuard() { for (int i = 0; i < 2; ++i) { cudaEventDestroy(a[i]); cudaEventDestroy(b[i]); } } } evg{ ev_perm, ev_done };

    auto gen_perm = [&](int buf, int64_t sweep0, int Sc) -> int {
        // wait until the sweep that last used this buffer has finished
        CUDA_TRY(cudaStreamWaitEvent(st2, ev_done[buf], 0));
        const int64_t nt = chains * Sc;
        if (perm16)
            perm_fill_kernel<uint16_t><<<(unsigned)nt, 256, 0, st2>>>(d_perm[buf].as<uint16_t>(), (int)chains, Sc, (int)p,
                                                                                 (int)sweep0, P.k0, P.k1, P.chain_offset);
        else
            perm_fill_kernel<int32_t><<<(unsigned)nt, 256, 0, st2>>>(d_perm[buf].as<int32_t>(), (int)chains, Sc, (int)p,
                                                                                (int)sweep0, P.k0, P.k1, P.chain_offset);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_perm[buf], st2));
        return BLIP_OK;
    };

    const int64_t nchunks = (n_total + S - 1) / S;
    const bool launch_log = getenv("BLIPGPU_LAUNCH_LOG") != nullptr;
    // permutation chunk boundaries: a short first chunk (nothing can hide its generation),
    // then PS sweeps each
    std::vector<int64_t> pstart;
    if (gen_perms) {
        pstart.push_back(0);
        int64_t at = std::min<int64_t>(S, n_total);
        while (at < n_total) { pstart.push_back(at); at = std::min<int64_t>(at + PS, n_total); }
        pstart.push_back(n_total);
    }
    const int64_t n_perm_chunks = !gen_perms ? 0 : (int64_t)pstart.size() - 1;
    if (gen_perms) BLIP_TRY(gen_perm(0, 0, (int)(pstart[1] - pstart[0])));
    int64_t pc = 0;
    int64_t max_chunk_use = 0;
    double max_per_sweep = 0.0;       // most values one kept sweep (all chains) has added to the arena
    unsigned long long used_before = 0;
    for (int64_t ch = 0; ch < nchunks; ++ch) {
        const int64_t sweep0 = ch * S;
        const int Sc = (int)std::min<int64_t>(S, n_total - sweep0);
        if (gen_perms) while (sweep0 >= pstart[(size_t)pc + 1]) ++pc;   // permutation chunk of this launch
        const int buf = !gen_perms ? 0 : (int)(pc & 1);
        const int64_t pc_sweep0 = !gen_perms ? sweep0 : pstart[(size_t)pc];
        const int64_t pc_end = !gen_perms ? sweep0 + Sc : pstart[(size_t)pc + 1];
        const int PSc = (int)(pc_end - pc_sweep0);
        if (inj_perm) BLIP_TRY(upload_chunk(d_perm[0].as<int32_t>(), streams->perm, chains, n_total * p, sweep0 * p, (int64_t)Sc * p, st));
        else if (sweep0 == pc_sweep0) {                             // first launch of a permutation chunk
            CUDA_TRY(cudaStreamWaitEvent(st, ev_perm[buf], 0));
            if (pc + 1 < n_perm_chunks)
                BLIP_TRY(gen_perm(buf ^ 1, pstart[(size_t)pc + 1], (int)(pstart[(size_t)pc + 2] - pstart[(size_t)pc + 1])));
        }
        if (inj_u) BLIP_TRY(upload_chunk(d_u.as<double>(), streams->u, chains, n_total * u_len, sweep0 * u_len, (int64_t)Sc * u_len, st));
        if (inj_z) BLIP_TRY(upload_chunk(d_zn.as<double>(), streams->z, chains, n_total * z_len, sweep0 * z_len, (int64_t)Sc * z_len, st));
        if (inj_ig) BLIP_TRY(upload_chunk(d_ig.as<double>(), streams->invgamma, chains, n_total, sweep0, (int64_t)Sc, st));
        if (inj_p0) BLIP_TRY(upload_chunk(d_p0.as<double>(), streams->p0_draw, chains, n_total * nprop, sweep0 * nprop, (int64_t)Sc * nprop, st));
        if (inj_g) BLIP_TRY(upload_chunk(d_g.as<double>(), streams->gamma_draw, chains, n_total, sweep0, (int64_t)Sc, st));

        // grow the value arena ahead of need (kept sweeps in this chunk only)
        const int64_t kept_in_chunk = std::max<int64_t>(0, std::min<int64_t>(sweep0 + Sc, n_total) - std::max<int64_t>(sweep0, burn));
        // Free space wanted before the launch: twice what a kept sweep has needed at most so far, or --
        // while nothing has been recorded yet -- a generous guess (p/16 non-zeros per row).  A launch that
        // still overflows is replayed below.
        int64_t need = max_per_sweep > 0 ? (int64_t)(2.0 * max_per_sweep * (double)kept_in_chunk) + chains * 64
                                         : chains * kept_in_chunk * std::max<int64_t>(64, p / 16);
        auto grow = [&](int64_t min_free) -> int {
            if (smp->arena_cap - (int64_t)used_before >= min_free) return BLIP_OK;
            // room for the rest of the run at the observed rate (+25 %), so that one growth is the norm
            const int64_t remaining_kept = n_total - std::max<int64_t>(sweep0, burn);
            int64_t newcap = (int64_t)used_before + std::max<int64_t>(min_free, (int64_t)(1.25 * max_per_sweep * (double)remaining_kept));
            double* nv = nullptr;
            const auto t_g0 = std::chrono::steady_clock::now();
            cudaError_t e = blip_store_alloc((void**)&nv, sizeof(double) * newcap, st);
            if (e != cudaSuccess) { blip_set_error("sample_spikeslab: cannot grow the value arena to %lld doubles: %s", (long long)newcap, cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
            CUDA_TRY(cudaMemcpyAsync(nv, smp->values, sizeof(double) * used_before, cudaMemcpyDeviceToDevice, st));
            blip_store_free(smp->values, st);
            smp->values = nv; smp->arena_cap = newcap;
            if (launch_log) fprintf(stderr, "[blipgpu] value arena grown to %lld doubles at sweep %lld (%.2f ms host clock)\n", (long long)newcap, (long long)sweep0,
                                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_g0).count());
            return BLIP_OK;
        };
        BLIP_TRY(grow(need));

        // snapshot the chain state so that an arena overflow can replay the chunk
        auto copy_state = [&](bool save) -> int {
            auto cp = [&](DevBuf& live, DevBuf& snap) -> int {
                if (!live.ptr) return BLIP_OK;
                CUDA_TRY(cudaMemcpyAsync(save ? snap.ptr : live.ptr, save ? live.ptr : snap.ptr, live.bytes, cudaMemcpyDeviceToDevice, st));
                return BLIP_OK;
            };
            BLIP_TRY(cp(d_beta, snap_beta)); BLIP_TRY(cp(d_h, snap_h)); BLIP_TRY(cp(d_mask, snap_mask));
            BLIP_TRY(cp(d_q, snap_q)); BLIP_TRY(cp(d_r, snap_r)); BLIP_TRY(cp(d_mu, snap_mu));
            BLIP_TRY(cp(d_ylat, snap_ylat)); BLIP_TRY(cp(d_ev, snap_ev));
            return BLIP_OK;
        };
        BLIP_TRY(copy_state(true));

        for (int attempt = 0;; ++attempt) {
            P.perm = (perm_in_kernel || block) ? nullptr : d_perm[buf].ptr; P.perm16 = perm16 ? 1 : 0;
            P.perm_S = PSc; P.perm_s0 = !gen_perms ? 0 : (int)(sweep0 - pc_sweep0);
            P.u = d_u.as<double>(); P.z = d_zn.as<double>(); P.invgamma = d_ig.as<double>();
            P.p0_draw = d_p0.as<double>(); P.gamma_draw = d_g.as<double>();
            P.sweep0 = (int)sweep0; P.S = Sc;
            P.values = smp->values; P.arena_cap = smp->arena_cap;
            CUDA_TRY(cudaMemsetAsync(d_ovf.ptr, 0, sizeof(int), st));
            unsigned grid = (unsigned)chains;
            if (gram && !block && !use_role) {
                // no more chains than resident CTAs: one CTA per chain, spread by the hardware
                P.sub_sweeps = (sub_sweeps_opt > 0 && chains > resident) ? std::min(sub_sweeps_opt, Sc) : Sc;
                const int64_t n_items = chains * ((Sc + P.sub_sweeps - 1) / P.sub_sweeps);
                grid = (unsigned)std::min<int64_t>(n_items, resident);
                CUDA_TRY(cudaMemsetAsync(d_queue.ptr, 0, d_queue.bytes, st));
            }
            {
                ScopedTimer timer(ctx, true);
                if (use_ws2c) {
                    // one 2-CTA cluster per chain: CTA 0 = decision warps + control warp, CTA 1 = helpers; q on both
                    cudaLaunchConfig_t lc;
                    memset(&lc, 0, sizeof(lc));
                    lc.gridDim = dim3(2 * grid); lc.blockDim = dim3(WS2_D_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    lc.attrs = at; lc.numAttrs = 1;
                    CUDA_TRY(cudaLaunchKernelEx(&lc, kern, P));
                } else if (use_cl || use_wbc) {
                    // one 2-CTA cluster per chain: CTA 0 decides, CTA 1 streams the Gram columns
                    cudaLaunchConfig_t lc;
                    memset(&lc, 0, sizeof(lc));
                    lc.gridDim = dim3(2 * grid); lc.blockDim = dim3(GB_BLOCK); lc.dynamicSmemBytes = smem; lc.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    lc.attrs = at; lc.numAttrs = 1;
                    CUDA_TRY(cudaLaunchKernelEx(&lc, kern, P));
                } else if (st_now) {
                    // one 2-CTA cluster per chain: CTA 0 resolves and keeps the books, CTA 1 owns q and streams
                    cudaLaunchConfig_t lc;
                    memset(&lc, 0, sizeof(lc));
                    lc.gridDim = dim3(2 * grid); lc.blockDim = dim3(ST_THREADS); lc.dynamicSmemBytes = smem_st; lc.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    lc.attrs = at; lc.numAttrs = 1;
                    CUDA_TRY(cudaLaunchKernelEx(&lc, kern_st, P));
                } else {
                    kern<<<grid, (use_ws || use_wb) ? WS_THREADS : GB_BLOCK, smem, st>>>(P);
                }
                CUDA_TRY(cudaGetLastError());
                const double ms = timer.stop(1);
                smp->sweep_ms += ms;
                smp->sweep_launches += 1;
                if (launch_log) fprintf(stderr, "[blipgpu] launch sweeps %lld..%lld grid %u: %.3f ms%s  (host clock since the call began: %.1f ms)\n",
                                        (long long)sweep0, (long long)(sweep0 + Sc), grid, ms, st_now ? " (stream)" : "",
                                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count());
            }
            int ovf = 0, nerr = 0;
            unsigned long long used = 0;
            CUDA_TRY(cudaMemcpyAsync(&ovf, d_ovf.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(&nerr, d_numerr.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(&used, smp->arena_used, sizeof(used), cudaMemcpyDeviceToHost, st));
            if (st_auto) {
                hs_now.resize((size_t)chains * GB_HS);
                CUDA_TRY(cudaMemcpyAsync(hs_now.data(), d_h.ptr, d_h.bytes, cudaMemcpyDeviceToHost, st));
            }
            CUDA_TRY(cudaStreamSynchronize(st));
            {
                if (nerr == 1) {
                    blip_set_error("dpotrf exited with INFO != 0: a block model's matrix is not positive definite. Try setting bsize=1.");
                    return BLIP_ERR_NUMERIC;
                }
                if (nerr) {
                    blip_set_error("sample_spikeslab: a chain's sigma2 / tau2 / p0 became non-finite in sweeps %lld..%lld "
                                   "(degenerate data or hyper-priors)", (long long)sweep0, (long long)(sweep0 + Sc));
                    return BLIP_ERR_NUMERIC;
                }
            }
            if (!ovf && st_auto) {
                double sum = 0.0, mx = 0.0;
                for (int64_t cc = 0; cc < chains; ++cc) {
                    const double d = hs_now[(size_t)(GB_HS * cc + 4)] - (hs_prev.empty() ? 0.0 : hs_prev[(size_t)(GB_HS * cc + 4)]);
                    sum += d; mx = std::max(mx, d);
                }
                hs_prev = hs_now;
                st_now = sum / (double)(chains * Sc) <= st_mean_max && mx / (double)Sc <= 1.5 * st_mean_max;
            }
            if (!ovf) {
                max_chunk_use = std::max<int64_t>(max_chunk_use, (int64_t)(used - used_before));
                if (kept_in_chunk > 0) max_per_sweep = std::max(max_per_sweep, (double)(used - used_before) / (double)kept_in_chunk);
                used_before = used;
                break;
            }
            if (attempt >= 8) { blip_set_error("sample_spikeslab: value arena overflow persists"); return BLIP_ERR_NOMEM; }
            // replay: restore state, rewind the arena, enlarge it to what this chunk asked for
            const int64_t wanted = (int64_t)(used - used_before);
            CUDA_TRY(cudaMemcpyAsync(smp->arena_used, &used_before, sizeof(used_before), cudaMemcpyHostToDevice, st));
            BLIP_TRY(copy_state(false));
            CUDA_TRY(cudaStreamSynchronize(st));
            max_chunk_use = std::max<int64_t>(max_chunk_use, wanted);
            if (kept_in_chunk > 0) max_per_sweep = std::max(max_per_sweep, (double)wanted / (double)kept_in_chunk);
            BLIP_TRY(grow(wanted + wanted / 4 + 1024));
        }
        if (gen_perms && sweep0 + Sc >= pc_end)
            CUDA_TRY(cudaEventRecord(ev_done[buf], st));           // last launch that reads this buffer
    }
    CUDA_TRY(cudaStreamSynchronize(st2));
    CUDA_TRY(cudaEventRecord(ev_call1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev_call0, ev_call1));
        smp->total_ms = ms;
        if (!hint_hit || (int64_t)used_before > ctx->arena_hint) {
            ctx->arena_hint = (int64_t)(1.15 * (double)used_before) + 1024;
            ctx->arena_hint_key[0] = chains; ctx->arena_hint_key[1] = n_keep; ctx->arena_hint_key[2] = p;
        }
        std::vector<double> hs((size_t)chains * GB_HS);
        CUDA_TRY(cudaMemcpy(hs.data(), d_h.ptr, d_h.bytes, cudaMemcpyDeviceToHost));
        for (int64_t c = 0; c < chains; ++c) {
            smp->n_changes += (int64_t)hs[(size_t)(GB_HS * c + 4)];
            smp->n_windows += (int64_t)hs[(size_t)(GB_HS * c + 5)];
        }
        // kernels: init + (x'y) + per chunk (permutation + sweep, sweeps counted by the timer)
        ctx->kernel_launches += 1 + (gram && design->kind == BLIPGPU_DESIGN_DENSE ? 1 : 0) + (gen_perms ? n_perm_chunks : 0);
        ctx->h2d_bytes += (int64_t)(probit ? sizeof(int32_t) : sizeof(double)) * n;
    }
    if (probit) {
        std::vector<uint32_t> ev((size_t)chains * 2);
        CUDA_TRY(cudaMemcpy(ev.data(), d_ev.ptr, d_ev.bytes, cudaMemcpyDeviceToHost));
        for (int64_t c = 0; c < chains; ++c) smp->n_events += ev[(size_t)(2 * c)];
    }
    guard.s = nullptr;
    *out = smp;
    return BLIP_OK;
}

extern "C" int blipgpu_sample_spikeslab(blipgpu_ctx* ctx, blipgpu_design* design,
                                        const blipgpu_spikeslab_config* cfg, const double* y,
                                        const int64_t* z, int32_t probit, int64_t chains,
                                        int64_t chain_offset, int64_t n_keep, int64_t burn,
                                        uint64_t seed, const blipgpu_streams* streams,
                                        const blipgpu_sample_options* opts, blipgpu_samples** out)
{
    return sample_impl(ctx, design, cfg, y, z, probit, chains, chain_offset, n_keep, burn, seed, streams,
                       opts, out, 1, 0, nullptr);
}

extern "C" int blipgpu_sample_spikeslab_multi(blipgpu_ctx* ctx, blipgpu_design* design,
                                              const blipgpu_spikeslab_config* cfg, int32_t bsize,
                                              int32_t max_signals_per_block, const double* y,
                                              const int64_t* z, int32_t probit, int64_t chains,
                                              int64_t chain_offset, int64_t n_keep, int64_t burn,
                                              uint64_t seed, const blipgpu_multi_streams* streams,
                                              const blipgpu_sample_options* opts, blipgpu_samples** out)
{
    if (bsize < 2) { blip_set_error("sample_spikeslab_multi: bsize must be at least 2 (use sample_spikeslab)"); return BLIP_ERR_ARG; }
    blipgpu_streams st;
    memset(&st, 0, sizeof(st));
    st.nprop = 1;
    if (streams) {
        st.u = streams->u; st.z = streams->z; st.invgamma = streams->invgamma;
        st.p0_draw = streams->p0_draw; st.gamma_draw = streams->gamma_draw; st.nprop = streams->nprop;
    }
    return sample_impl(ctx, design, cfg, y, z, probit, chains, chain_offset, n_keep, burn, seed,
                       streams ? &st : nullptr, opts, out, bsize, max_signals_per_block,
                       streams ? streams->blocks : nullptr);
}
