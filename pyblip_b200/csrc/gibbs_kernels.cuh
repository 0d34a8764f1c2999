// gibbs_kernels.cuh -- device side of the multi-chain single-site spike-and-slab
// Gibbs sampler (linear + probit), sm_100a.
//
// Semantics follow /root/reference/pyblip/linear/_linear.pyx:126-229 (coordinate
// sweep) and pyblip/cython_utils/_update_hparams.pyx:51-91 (hyper-parameters),
// re-designed around EXACT SPECULATION: a null coordinate (beta_j == 0 before and
// after its visit) leaves the chain state untouched, so a whole window of
// upcoming coordinates is evaluated in parallel against the current state and
// only the first coordinate whose value changes is applied; the window restarts
// right after it.  The visiting order, the variates and every decision are
// identical to a strictly sequential sweep.
//
// Two formulations:
//   residual  state r = y - X beta in shared memory; a window is 32 columns of X
//             streamed from HBM/L2 (one 8n-byte column per update); required for
//             probit, where every change redraws all n latents.
//   gram      state q = X^T r in shared memory; a window is one coordinate per
//             thread at O(1) cost and a change reads one 8p-byte Gram column.
#pragma once
#include "internal.cuh"
#include <cstdio>

#ifndef BLIP_GB_BLOCK
#define BLIP_GB_BLOCK 256
#endif
constexpr int GB_BLOCK = BLIP_GB_BLOCK;   // (tuning variants only: the residual and role-specialised kernels assume 256)
constexpr int GB_NWARP = GB_BLOCK / 32;
// x modulo the block size, x >= -GB_BLOCK (the sliding window's position <-> thread map)
__device__ __forceinline__ int gb_mod(int x)
{
    if ((GB_BLOCK & (GB_BLOCK - 1)) == 0) return x & (GB_BLOCK - 1);
    return (x + GB_BLOCK) % GB_BLOCK;
}
constexpr int GB_ROUND = 32;      // residual mode: columns speculated per window (64 with two scoring warps was
                                  // measured: no gain on probit C3, -20 % on linear C2 -- more columns redone after a change)
constexpr int GB_HS = 6;          // doubles of per-chain scalar state: sigma2 tau2 p0 rr n_changes n_windows

struct SweepParams {
    // design
    const double* XT; int64_t ldx; const double* Xl2; const double* G; int64_t ldg;
    int n, p, words, n2, p2;
    // data
    const double* q0; double yy; const int32_t* zlab;
    // model
    blipgpu_spikeslab_config cfg;
    // chain state (global, persists between launches)
    double* beta;      // [chains][p]
    double* r;         // [chains][n2]            residual mode
    double* mu;        // [chains][n2]            probit
    double* ylat;      // [chains][n2]            probit
    double* q;         // [chains][p2]            gram mode
    double* hstate;    // [chains][GB_HS]  sigma2, tau2, p0, rr, #changes, #windows
    uint32_t* mask;    // [chains][words]  active-set bitmask
    uint32_t* events;  // [chains][2]  probit: redraw events so far, latents-dirty flag
    // random streams
    uint32_t k0, k1; uint32_t chain_offset;
    const void* perm;          // [chains][perm_S][p] uint16 or int32 (generated or injected);
    int perm16, perm_S, perm_s0; //   this launch's sweep s is row perm_s0 + s of a chain
    const double* u;           // [chains][S][p] or NULL
    const double* z;           // [chains][S][p] or NULL
    const double* invgamma;    // [chains][S] or NULL
    const double* p0_draw;     // [chains][S][nprop] or NULL
    const double* gamma_draw;  // [chains][S] or NULL
    int nprop;
    // sweep range of this launch
    int sweep0, S, burn, n_keep;
    // outputs
    uint32_t* bits; int64_t rows, rows_pad;
    int64_t* row_off; int32_t* row_nnz;
    double* values; int64_t arena_cap; unsigned long long* arena_used; int* overflow;
    double* hyper_out; double* latents;
    int gram_refresh;
    // block sampler (bsize > 1, gibbs_block.cuh)
    const int32_t* blocks;     // [chains][nblock][bsize] block table in visiting order
    const double* gb;          // [nblock][bsize*bsize]   X_B^T X_B of the original block b = first member / bsize
    const uint32_t* model_masks; // [nmodel-1] member bit sets of the non-empty models, reference order
    int bsize, nblock, nmodel, npad, blocks_per_warp; int* numeric_err;   // npad: nmodel rounded up to a power of two
    // gram mode work queue: items (chain, sub-chunk of sub_sweeps sweeps) handed out in order;
    // progress[c] = sub-chunks of chain c completed in this launch
    int* queue; int* progress; int n_chains, sub_sweeps;
    int64_t scr_stride;        // warp-specialised gram kernel: elements between the two halves of scrP / scrL
    void* scrP;                // gram mode, keyed visiting order: [chains][p] uint16/int32, rewritten every sweep
    float* scrL;               // gram mode: [chains][p]  logit of the spike uniform per position (NaN = exact path)
    double2* scrZ;             // gram mode: [chains][p]  (sqrt(post_var) * z, post_var) at positions of active coordinates
    int* scrI;                 // column-stream gram kernel: [2][chains][p2]  position of every coordinate in the visiting order
};

struct PermView {           // visiting order of one (chain, sweep)
    const void* base; int64_t off; int is16;
    __device__ __forceinline__ int operator[](int pos) const {   // L2 loads: the gram kernel's
        return is16 ? (int)__ldcg(reinterpret_cast<const uint16_t*>(base) + off + pos)   // scratch copy is
                    : __ldcg(reinterpret_cast<const int32_t*>(base) + off + pos);         // rewritten every sweep
    }
};
__device__ __forceinline__ PermView perm_view(const SweepParams& P, int c, int s)
{
    return PermView{ P.perm, ((int64_t)c * P.perm_S + P.perm_s0 + s) * P.p, P.perm16 };
}

// ---------------------------------------------------------------------------
// the scalar part of one coordinate update (_linear.pyx:146-158), fp64, with the
// reference's association order.  Compiled with -fmad=false so nothing fuses.
// ---------------------------------------------------------------------------
struct Hyper { double sigma2, tau2, logodds; };

__device__ __forceinline__ double kappa_of(double XjTr, double xl2, const Hyper& h)
{
    double logdet = log(1.0 + h.tau2 * xl2 / h.sigma2) / 2.0;          // :129
    double log_ratio_num = h.tau2 / h.sigma2 * XjTr * XjTr;            // :146
    double log_ratio_denom = 2 * (h.sigma2 + h.tau2 * xl2);            // :147
    double ratio = exp(h.logodds - log_ratio_num / log_ratio_denom + logdet);
    return ratio / (1.0 + ratio);                                      // :149
}

__device__ __forceinline__ double slab_draw(double XjTr, double xl2, const Hyper& h, double znorm)
{
    double post_var = 1.0 / (1.0 / h.tau2 + xl2 / h.sigma2);           // :130
    double post_mean = post_var * XjTr / h.sigma2;                     // :155
    return sqrt(post_var) * znorm + post_mean;                         // :158
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix,
// *total gets the block sum.  s_warp: >= GB_NWARP+1 ints of shared memory.
template <class SY = SyncCta>
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total)
{
    const int lane = SY::tid() & 31, warp = SY::tid() >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    SY::sync();
    if (warp == 0) {
        int w = lane < GB_NWARP ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < GB_NWARP; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < GB_NWARP) s_warp[lane] = wi - w;       // exclusive warp offsets
        if (lane == GB_NWARP - 1) s_warp[GB_NWARP] = wi;  // total
    }
    SY::sync();
    int res = s_warp[warp] + incl - v;
    *total = s_warp[GB_NWARP];
    SY::sync();
    return res;
}

// ---------------------------------------------------------------------------
// end-of-sweep: hyper-parameter draws (_update_hparams.pyx:51-91) and recording
// ---------------------------------------------------------------------------
struct SweepShared {
    double sigma2, tau2, p0, rr, logodds;
    long long off;
    int scan[GB_NWARP + 2];
};

// Called by every thread of the block (or of the sub-group SY synchronises: GB_BLOCK threads with
// threadIdx.x < GB_BLOCK).  ssr = ||r||^2 (already reduced, same value in every thread).
// Updates sh->{sigma2,tau2,p0}; records the sweep if kept.
template <class SY = SyncCta>
__device__ inline void sweep_epilogue(const SweepParams& P, SweepShared* sh, const uint32_t* mask_s,
                                      const double* beta_g, double* scratch, const RngKey& key,
                                      int c, int s, double ssr, const double* y_s, bool latents_dirty)
{
    const int tid = SY::tid();
    const int sweep = P.sweep0 + s;
    // number of active coordinates and sum of squares (_update_hparams.pyx:52-54, 85-88)
    int cnt = 0;
    double ss = 0.0;
    for (int w = tid; w < P.words; w += GB_BLOCK) {
        uint32_t m = mask_s[w];
        cnt += __popc(m);
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            double v = __ldcg(beta_g + 32 * w + b);
            ss = fma(v, v, ss);
        }
    }
    ss = block_sum<SY>(ss, scratch);
    const int k = (int)(block_sum<SY>((double)cnt, scratch) + 0.5);

    if (tid == 0) {
        const blipgpu_spikeslab_config& cf = P.cfg;
        const int64_t hs = (int64_t)c * P.S + s;
        if (cf.update_p0 == 1) {
            double a = cf.p0_a0 + P.p - k, b = cf.p0_b0 + k;
            if (cf.min_p0 == 0) {
                sh->p0 = P.p0_draw ? P.p0_draw[hs * P.nprop] : rng_beta(key, 0, sweep, a, b);
            } else {
                double p0 = cf.min_p0;
                for (int jj = 0; jj < 100; ++jj) {
                    double prop = P.p0_draw ? P.p0_draw[hs * P.nprop + jj]
                                            : rng_beta(key, jj, sweep, a, b);
                    if (prop > cf.min_p0) { p0 = prop; break; }
                }
                sh->p0 = p0;
            }
        }
        if (cf.update_sigma2 == 1) {
            // gram form tracks ||r||^2 by a recurrence (and rebuilds it as y'y - b'(q0+q)); on a
            // near-exact fit both cancel and can land a rounding error below zero, where the
            // reference's dnrm2(r) cannot: clamp, or sqrt would poison sigma2 and every kappa after it
            ssr = fmax(ssr, 0.0);
            double r2 = sqrt(ssr);             // dnrm2 ...
            r2 = r2 * r2;                      // ... squared (:77-78)
            double sigma_b = r2 / 2.0 + cf.sigma2_b0;
            double ig = P.invgamma ? P.invgamma[hs]
                                   : 1.0 / rng_gamma(key, SLOT_INVGAMMA, sweep, P.n / 2.0 + cf.sigma2_a0);
            sh->sigma2 = sigma_b * ig;
        }
        if (cf.update_tau2) {
            double g = P.gamma_draw ? P.gamma_draw[hs]
                                    : rng_gamma(key, SLOT_TAU_GAMMA, sweep, cf.tau2_a0 + (double)k / 2.0);
            sh->tau2 = (cf.tau2_b0 + ss / 2.0) / g;
        }
        // a non-finite hyper-parameter makes every later decision meaningless: report it (the host
        // raises BLIPGPU_ERR_NUMERIC) instead of recording garbage.  p0 == 1 is legal (quirk 1).
        if (!isfinite(sh->sigma2) || !isfinite(sh->tau2) || isnan(sh->p0)) atomicMax(P.numeric_err, 2);
    }
    SY::sync();

    if (sweep < P.burn) return;               // uniform across the block
    const int64_t row = (int64_t)c * P.n_keep + (sweep - P.burn);
    if (tid == 0) {
        P.hyper_out[row] = sh->p0;
        P.hyper_out[P.rows + row] = sh->tau2;
        P.hyper_out[2 * P.rows + row] = sh->sigma2;
        unsigned long long off = atomicAdd(P.arena_used, (unsigned long long)k);
        long long o = (long long)off;
        if (off + (unsigned long long)k > (unsigned long long)P.arena_cap) { *P.overflow = 1; o = -1; }
        sh->off = o;
        P.row_off[row] = o;
        P.row_nnz[row] = k;
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) P.bits[(int64_t)w * P.rows_pad + row] = mask_s[w];
    SY::sync();
    const long long off = sh->off;
    if (off >= 0 && k > 0) {
        int base = 0;
        for (int w0 = 0; w0 < P.words; w0 += GB_BLOCK) {     // uniform trip count
            int w = w0 + tid;
            uint32_t m = (w < P.words) ? mask_s[w] : 0u;
            int total;
            int pre = block_excl_scan<SY>(__popc(m), sh->scan, &total);
            double* dst = P.values + off + base + pre;
            while (m) {
                int b = __ffs(m) - 1;
                m &= m - 1;
                *dst++ = __ldcg(beta_g + 32 * w + b);
            }
            base += total;
        }
    }
    if (P.latents && latents_dirty && y_s) {
        double* dst = P.latents + row * (int64_t)P.n;
        for (int k2 = tid; k2 < P.n; k2 += GB_BLOCK) dst[k2] = y_s[k2];
    }
}

// ===========================================================================
// residual mode
// ===========================================================================
struct ResidualBcast { int tstar, jstar; double beta_old, dspec, beta_new; };

template <bool PROBIT>
__global__ void __launch_bounds__(GB_BLOCK, 2) gibbs_residual_kernel(SweepParams P)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SweepShared sh;
    __shared__ ResidualBcast bc;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.n, p = P.p, n2 = P.n2;
    double* r_s = sm;
    double* mu_s = PROBIT ? r_s + n2 : nullptr;
    double* y_s = PROBIT ? r_s + 2 * n2 : nullptr;
    double* part = r_s + (PROBIT ? 3 : 1) * n2;        // [2][GB_ROUND] dots of the current and of the next window (+ slack)
    double* scratch = part + GB_NWARP * 32;            // [40]
    uint32_t* mask_s = (uint32_t*)(scratch + 40);      // [words]

    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    double* beta_g = P.beta + (int64_t)c * p;
    double* r_g = P.r + (int64_t)c * n2;

    for (int k = tid; k < n2; k += GB_BLOCK) {
        r_s[k] = r_g[k];
        if (PROBIT) { mu_s[k] = P.mu[(int64_t)c * n2 + k]; y_s[k] = P.ylat[(int64_t)c * n2 + k]; }
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) mask_s[w] = P.mask[(int64_t)c * P.words + w];
    if (tid == 0) {
        sh.sigma2 = P.hstate[GB_HS * c + 0]; sh.tau2 = P.hstate[GB_HS * c + 1];
        sh.p0 = P.hstate[GB_HS * c + 2]; sh.rr = 0.0;
    }
    long long n_changes = 0, n_windows = 0;     // block-uniform work counters
    uint32_t event = PROBIT ? P.events[2 * c] : 0u;
    bool dirty = PROBIT ? (P.events[2 * c + 1] != 0u) : false;
    __syncthreads();

    for (int s = 0; s < P.S; ++s) {
        const int sweep = P.sweep0 + s;
        const int64_t soff = ((int64_t)c * P.S + s) * p;
        const PermView perm_s = perm_view(P, c, s);
        const double* u_s = P.u ? P.u + soff : nullptr;
        const double* z_s = P.z ? P.z + soff : nullptr;
        if (tid == 0) sh.logodds = log(sh.p0) - log(1 - sh.p0);   // :76, :222
        __syncthreads();
        const Hyper h = { sh.sigma2, sh.tau2, sh.logodds };

        // Two-stage pipeline over the speculation windows.  While warp 0 scores window w (dependent fp64
        // math, ~1 us), warps 1-7 already stream their columns of window w+1 against the same residual
        // (warp 0 streams its own four afterwards) -- valid if w turns out to be all null, which is the
        // common case; a change in w discards them.
        // part[cur] holds the dots of the window at pos0 when have_dots is set.  A warp always streams
        // whole columns with the same lane -> element assignment, so a dot is the same number whichever
        // warp computes it: results do not depend on the schedule.
        int pos0 = 0, cur = 0;
        bool have_dots = false;
        int null_run = 0;                              // consecutive all-null windows so far
        while (pos0 < p) {
            const int T = min(GB_ROUND, p - pos0);
            ++n_windows;
            // Measured at BASELINE's shapes: the pipeline below gives linear C2 (n = 1000) +25 % in the stationary
            // regime; probit C3 (n = 2000; a change -- hence a discarded 512 KB window -- in 15 % of the windows,
            // scoring only 5 % of a window's time) loses 7 % with it, and 9 % even with speculation switched off
            // inside the restructured loop: probit keeps the plain schedule of round 1, verbatim.
            // The linear sweep uses it too until two windows in a row were null: from beta = 0 nearly every
            // window holds a change, and every speculative window would be thrown away.
            constexpr bool PIPE = !PROBIT;
            if (!PIPE || null_run < 2) {
                // ---- speculative dots of the next T columns against the current residual
                const int jl = perm_s[min(pos0 + lane, p - 1)];
                // warp 0 scores the window after the dots: what does not depend on them (indicator, uniform, ||X_j||^2)
                // is fetched now, under the latency of the column loads
                bool act = false;
                double u_spec = 0.0, xl2_spec = 0.0;
                if (warp == 0 && lane < T) {
                    act = (mask_s[jl >> 5] >> (jl & 31)) & 1u;
                    if (!act) {
                        u_spec = u_s ? u_s[pos0 + lane] : rng_uniform(key, KIND_SPIKE_U, pos0 + lane, sweep, 0);
                        xl2_spec = P.Xl2[jl];
                    }
                }
                // each warp owns GB_ROUND / GB_NWARP whole columns: few accumulators, so many loads stay in flight
                constexpr int CPW = GB_ROUND / GB_NWARP;
                const double* xc[CPW];
                double acc[CPW];
    #pragma unroll
                for (int i = 0; i < CPW; ++i) {
                    xc[i] = P.XT + (int64_t)__shfl_sync(0xffffffffu, jl, warp * CPW + i) * P.ldx;
                    acc[i] = 0.0;
                }
    #pragma unroll 4
                for (int k = 2 * lane; k < n2; k += 64) {
                    const double2 rk = *reinterpret_cast<const double2*>(r_s + k);
    #pragma unroll
                    for (int i = 0; i < CPW; ++i) {
                        const double2 x = __ldg(reinterpret_cast<const double2*>(xc[i] + k));
                        acc[i] = fma(x.x, rk.x, acc[i]);
                        acc[i] = fma(x.y, rk.y, acc[i]);
                    }
                }
    #pragma unroll
                for (int i = 0; i < CPW; ++i) {
                    const double t = warp_sum(acc[i]);
                    if (lane == i) part[warp * CPW + i] = t;         // part[column of the window]
                }
                __syncthreads();
                if (warp == 0) {
                    const double d = part[lane];
                    bool change = false;
                    if (lane < T) {
                        if (act) change = true;
                        else change = !(u_spec <= kappa_of(d, xl2_spec, h));   // NaN kappa -> slab, like the reference
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, change);
                    const int tstar = m ? __ffs(m) - 1 : -1;
                    if (lane == (tstar < 0 ? 0 : tstar)) {
                        bc.tstar = tstar; bc.jstar = jl;
                        bc.beta_old = act ? __ldcg(beta_g + jl) : 0.0;
                        bc.dspec = d;
                    }
                }
                __syncthreads();
                const int tstar_plain = bc.tstar;
                if (tstar_plain < 0) { pos0 += T; ++null_run; have_dots = false; continue; }
                null_run = 0;
            } else {
                bool spec_skip = false;                        // (warp 0 only)
                // what the scoring of window pos0 needs besides the dots (warp 0): indicator, uniform, ||X_j||^2
                int jl0 = 0;
                bool act = false;
                double u_spec = 0.0, xl2_spec = 0.0;
                auto score_inputs = [&]() {
                    jl0 = perm_s[min(pos0 + lane, p - 1)];
                    if (lane < T) {
                        act = (mask_s[jl0 >> 5] >> (jl0 & 31)) & 1u;
                        if (!act) {
                            u_spec = u_s ? u_s[pos0 + lane] : rng_uniform(key, KIND_SPIKE_U, pos0 + lane, sweep, 0);
                            xl2_spec = P.Xl2[jl0];
                        }
                    }
                };
                const bool had_dots = have_dots;
                if (!have_dots) {
                    // ---- first window of a sweep, or the one after a change: all eight warps, four columns each;
                    // warp 0 fetches its scoring inputs first, under the latency of the column loads
                    if (warp == 0) score_inputs();
                    const int jl = perm_s[min(pos0 + lane, p - 1)];
                    constexpr int CPW = GB_ROUND / GB_NWARP;
                    const double* xc[CPW];
                    double acc[CPW];
    #pragma unroll
                    for (int i = 0; i < CPW; ++i) {
                        xc[i] = P.XT + (int64_t)__shfl_sync(0xffffffffu, jl, warp * CPW + i) * P.ldx;
                        acc[i] = 0.0;
                    }
    #pragma unroll 4
                    for (int k = 2 * lane; k < n2; k += 64) {
                        const double2 rk = *reinterpret_cast<const double2*>(r_s + k);
    #pragma unroll
                        for (int i = 0; i < CPW; ++i) {
                            const double2 x = __ldg(reinterpret_cast<const double2*>(xc[i] + k));
                            acc[i] = fma(x.x, rk.x, acc[i]);
                            acc[i] = fma(x.y, rk.y, acc[i]);
                        }
                    }
    #pragma unroll
                    for (int i = 0; i < CPW; ++i) {
                        const double t = warp_sum(acc[i]);
                        if (lane == i) part[cur * GB_ROUND + warp * CPW + i] = t;   // part[column of the window]
                    }
                    __syncthreads();
                }
                const int pos1 = pos0 + T;                     // start of the next window
                const bool has_next = pos1 < p;             // (this branch runs only after two null windows)
                if (warp == 0) {
                    // ---- score window pos0 (its dots came from the previous iteration's speculation: the other
                    // warps are already streaming the next window, so the inputs' latency hides behind them)
                    if (had_dots) score_inputs();
                    const int jl = jl0;
                    const double d = part[cur * GB_ROUND + lane];
                    bool change = false;
                    if (lane < T) {
                        if (act) change = true;
                        else change = !(u_spec <= kappa_of(d, xl2_spec, h));   // NaN kappa -> slab, like the reference
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, change);
                    const int tstar = m ? __ffs(m) - 1 : -1;
                    if (lane == (tstar < 0 ? 0 : tstar)) {
                        bc.tstar = tstar; bc.jstar = jl;
                        bc.beta_old = act ? __ldcg(beta_g + jl) : 0.0;
                        bc.dspec = d;
                    }
                    spec_skip = tstar >= 0;                    // a change: the next window's dots would be discarded
                }
                if (has_next && !(warp == 0 && spec_skip)) {
                    // ---- meanwhile (warp 0: afterwards): the dots of window pos1, four columns per warp; the other
                    // seven warps' loads are in flight while warp 0 scores, and warp 0's own four columns stream
                    // at the bandwidth the others leave free
                    const int jl = perm_s[min(pos1 + lane, p - 1)];
                    constexpr int CPW = GB_ROUND / GB_NWARP;
                    const double* xc[CPW];
                    double acc[CPW];
    #pragma unroll
                    for (int i = 0; i < CPW; ++i) {
                        xc[i] = P.XT + (int64_t)__shfl_sync(0xffffffffu, jl, warp * CPW + i) * P.ldx;
                        acc[i] = 0.0;
                    }
    #pragma unroll 4
                    for (int k = 2 * lane; k < n2; k += 64) {
                        const double2 rk = *reinterpret_cast<const double2*>(r_s + k);
    #pragma unroll
                        for (int i = 0; i < CPW; ++i) {
                            const double2 x = __ldg(reinterpret_cast<const double2*>(xc[i] + k));
                            acc[i] = fma(x.x, rk.x, acc[i]);
                            acc[i] = fma(x.y, rk.y, acc[i]);
                        }
                    }
    #pragma unroll
                    for (int i = 0; i < CPW; ++i) {
                        const double t = warp_sum(acc[i]);
                        if (lane == i) part[(cur ^ 1) * GB_ROUND + warp * CPW + i] = t;
                    }
                }
                __syncthreads();
                if (bc.tstar < 0) { pos0 = pos1; cur ^= 1; have_dots = has_next; ++null_run; continue; }
                have_dots = false;
                null_run = 0;
            }
            const int tstar = bc.tstar;

            // ---- apply the first changing coordinate exactly like the reference
            ++n_changes;
            const int j = bc.jstar;
            const double beta_old = bc.beta_old;
            const double* xcol = P.XT + (int64_t)j * P.ldx;
            double d = bc.dspec;
            if (beta_old != 0) {                      // :136-145 add back, fresh dot
                double partial = 0.0;
                for (int k = 2 * tid; k < n2; k += 2 * GB_BLOCK) {
                    double2 x = __ldg(reinterpret_cast<const double2*>(xcol + k));
                    double2 rk = *reinterpret_cast<double2*>(r_s + k);
                    rk.x = fma(beta_old, x.x, rk.x); rk.y = fma(beta_old, x.y, rk.y);
                    *reinterpret_cast<double2*>(r_s + k) = rk;
                    partial = fma(x.x, rk.x, partial); partial = fma(x.y, rk.y, partial);
                }
                d = block_sum(partial, scratch);
            }
            if (tid == 0) {
                const int pos = pos0 + tstar;
                double u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                double xl2 = P.Xl2[j];
                double kappa = kappa_of(d, xl2, h);
                double bnew = 0.0;
                if (!(u <= kappa)) {
                    double zn = z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0);
                    bnew = slab_draw(d, xl2, h, zn);
                }
                bc.beta_new = bnew;
                __stcg(beta_g + j, bnew);
                if (bnew != 0) mask_s[j >> 5] |= (1u << (j & 31));
                else mask_s[j >> 5] &= ~(1u << (j & 31));
            }
            __syncthreads();
            const double bnew = bc.beta_new;
            if (!PROBIT) {
                if (bnew != 0) {                      // :160-170
                    for (int k = 2 * tid; k < n2; k += 2 * GB_BLOCK) {
                        double2 x = __ldg(reinterpret_cast<const double2*>(xcol + k));
                        double2 rk = *reinterpret_cast<double2*>(r_s + k);
                        rk.x = fma(-bnew, x.x, rk.x); rk.y = fma(-bnew, x.y, rk.y);
                        *reinterpret_cast<double2*>(r_s + k) = rk;
                    }
                }
            } else {
                const double delta = bnew - beta_old; // :172-190
                if (delta != 0) {
                    for (int k = tid; k < n; k += GB_BLOCK) {
                        // mu += delta * X_j with a rounded product, not an fma: when a coefficient goes
                        // back to zero its share of mu then cancels to exactly 0.0 (an fma would leave the
                        // rounding residual of the product, of either sign), and truncnorm_std branches on
                        // the sign of a = -mu/scale at a == 0 (_truncnorm.pyx:57-69)
                        double m = __dadd_rn(__dmul_rn(delta, __ldg(xcol + k)), mu_s[k]);
                        mu_s[k] = m;
                        double yv = truncnorm(key, KIND_TRUNCNORM, (uint32_t)k, event, m, P.cfg.sigma2,
                                              0.0, P.zlab[k]);
                        y_s[k] = yv;
                        r_s[k] = yv - m;
                    }
                    event++;
                    dirty = true;
                }
            }
            __syncthreads();
            pos0 += tstar + 1;
        }

        // ---- hyper-parameters + recording
        double ssr = 0.0;
        for (int k = tid; k < n; k += GB_BLOCK) ssr = fma(r_s[k], r_s[k], ssr);
        ssr = block_sum(ssr, scratch);
        sweep_epilogue(P, &sh, mask_s, beta_g, scratch, key, c, s, ssr, y_s, dirty);
        dirty = false;
        __syncthreads();
    }

    for (int k = tid; k < n2; k += GB_BLOCK) {
        r_g[k] = r_s[k];
        if (PROBIT) { P.mu[(int64_t)c * n2 + k] = mu_s[k]; P.ylat[(int64_t)c * n2 + k] = y_s[k]; }
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2; P.hstate[GB_HS * c + 2] = sh.p0;
        P.hstate[GB_HS * c + 4] += (double)n_changes; P.hstate[GB_HS * c + 5] += (double)n_windows;
        if (PROBIT) { P.events[2 * c] = event; P.events[2 * c + 1] = dirty ? 1u : 0u; }
    }
}

// ===========================================================================
// gram mode (linear only)
// ===========================================================================
struct GramBcast { double delta; };
#ifdef BLIP_PHASE_CLOCKS
#define PH_DECL long long ph_t[12] = {0,0,0,0,0,0,0,0,0,0,0,0}; long long ph_last = clock64();
#define PH(i) do { if (threadIdx.x == 0) { long long _n = clock64(); ph_t[i] += _n - ph_last; ph_last = _n; } } while (0)
#else
#define PH_DECL
#define PH(i) do {} while (0)
#endif
#ifndef BLIP_GB_AXU
#define BLIP_GB_AXU 10
#endif
constexpr int GB_AXU = BLIP_GB_AXU;    // 16-byte Gram loads a thread keeps in flight
#ifndef BLIP_GB_PEND
#define BLIP_GB_PEND 2
#endif
constexpr int GB_PEND = BLIP_GB_PEND;   // changes whose Gram columns are applied to q in one pass
constexpr int GB_FLUSH_SLOTS = GB_AXU / GB_PEND;
#ifndef BLIP_GB_GRAM_REGS
#define BLIP_GB_GRAM_REGS 128
#endif
constexpr int GB_GRAM_REGS = BLIP_GB_GRAM_REGS;  // 2 CTAs x 256 threads x 128 registers = the whole register file

template <int DESIGN> __device__ __forceinline__ double2 gram_pair(const SweepParams& P, int jstar, int k)
{
    if (DESIGN == BLIPGPU_DESIGN_CHANGEPOINT) {
        // X[i][j] = 1{i >= j}  =>  (X^T X)[a][b] = T - max(a, b)
        return make_double2(k < P.p ? (double)(P.n - max(jstar, k)) : 0.0,
                            k + 1 < P.p ? (double)(P.n - max(jstar, k + 1)) : 0.0);
    } else {
        return __ldg(reinterpret_cast<const double2*>(P.G + (int64_t)jstar * P.ldg + k));
    }
}
template <int DESIGN> __device__ __forceinline__ double gram_entry(const SweepParams& P, int a, int b)
{
    if (DESIGN == BLIPGPU_DESIGN_CHANGEPOINT) return (double)(P.n - max(a, b));
    else return __ldg(P.G + (int64_t)a * P.ldg + b);
}

// q -= delta * G[:, jstar], the whole block; GB_AXU independent 16-byte loads are
// issued before the first use so that one CTA keeps ~40 KB of the column in flight.
template <int DESIGN>
__device__ __forceinline__ void gram_axpy(const SweepParams& P, double* q, int jstar, double delta)
{
    const int p2 = P.p2;
    for (int k0 = 2 * threadIdx.x; k0 < p2; k0 += 2 * GB_BLOCK * GB_AXU) {
        double2 g[GB_AXU];
#pragma unroll
        for (int u = 0; u < GB_AXU; ++u) {
            const int k = k0 + u * 2 * GB_BLOCK;
            g[u] = (k < p2) ? gram_pair<DESIGN>(P, jstar, k) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < GB_AXU; ++u) {
            const int k = k0 + u * 2 * GB_BLOCK;
            if (k < p2) {
                double2 qq = *reinterpret_cast<double2*>(q + k);
                qq.x = fma(-delta, g[u].x, qq.x); qq.y = fma(-delta, g[u].y, qq.y);
                *reinterpret_cast<double2*>(q + k) = qq;
            }
        }
    }
}

// q -= sum_k pd[k] * G[:, pj[k]] for the npend pending changes, applied element-wise in
// the order the changes happened (the same fma sequence as npend separate passes), but
// with q read and written once.  Block-wide; the caller synchronises.
template <int DESIGN>
__device__ __forceinline__ void gram_flush(const SweepParams& P, double* q, const int* pj_s,
                                           const double* pd_s, int npend)
{
    const int p2 = P.p2;
    int pj[GB_PEND]; double pd[GB_PEND];
#pragma unroll
    for (int k = 0; k < GB_PEND; ++k) { pj[k] = k < npend ? pj_s[k] : 0; pd[k] = k < npend ? pd_s[k] : 0.0; }
    int k0 = 2 * threadIdx.x;
    if (DESIGN == BLIPGPU_DESIGN_DENSE && npend == GB_PEND) {
        // the common case, a full list on a stored Gram matrix: per-thread column pointers with
        // immediate offsets and no bounds tests while every thread of the block is in range
        const double2* col[GB_PEND];
#pragma unroll
        for (int k = 0; k < GB_PEND; ++k)
            col[k] = reinterpret_cast<const double2*>(P.G + (int64_t)pj[k] * P.ldg) + threadIdx.x;
        double2* qv = reinterpret_cast<double2*>(q) + threadIdx.x;
        const int nfull = (p2 / 2) / GB_BLOCK;             // slots in which all 256 threads are in range
        int slot = 0;
        for (; slot + GB_FLUSH_SLOTS <= nfull; slot += GB_FLUSH_SLOTS) {
            double2 g[GB_FLUSH_SLOTS][GB_PEND];
#pragma unroll
            for (int u = 0; u < GB_FLUSH_SLOTS; ++u)
#pragma unroll
                for (int k = 0; k < GB_PEND; ++k) g[u][k] = __ldg(col[k] + (slot + u) * GB_BLOCK);
#pragma unroll
            for (int u = 0; u < GB_FLUSH_SLOTS; ++u) {
                double2 qq = qv[(slot + u) * GB_BLOCK];
#pragma unroll
                for (int k = 0; k < GB_PEND; ++k) { qq.x = fma(-pd[k], g[u][k].x, qq.x); qq.y = fma(-pd[k], g[u][k].y, qq.y); }
                qv[(slot + u) * GB_BLOCK] = qq;
            }
        }
        k0 += slot * 2 * GB_BLOCK;
    }
    for (; k0 < p2; k0 += 2 * GB_BLOCK * GB_FLUSH_SLOTS) {
        double2 g[GB_FLUSH_SLOTS][GB_PEND];
#pragma unroll
        for (int u = 0; u < GB_FLUSH_SLOTS; ++u) {
            const int kk = k0 + u * 2 * GB_BLOCK;
#pragma unroll
            for (int k = 0; k < GB_PEND; ++k)
                g[u][k] = (kk < p2 && k < npend) ? gram_pair<DESIGN>(P, pj[k], kk) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < GB_FLUSH_SLOTS; ++u) {
            const int kk = k0 + u * 2 * GB_BLOCK;
            if (kk < p2) {
                double2 qq = *reinterpret_cast<double2*>(q + kk);
#pragma unroll
                for (int k = 0; k < GB_PEND; ++k)
                    if (k < npend) { qq.x = fma(-pd[k], g[u][k].x, qq.x); qq.y = fma(-pd[k], g[u][k].y, qq.y); }
                *reinterpret_cast<double2*>(q + kk) = qq;
            }
        }
    }
}

// The state-independent part of the slab draw of an active coordinate (visited at `pos`):
// (sqrt(post_var) * z, post_var), _linear.pyx:130,158.
__device__ __forceinline__ void slab_prepare(const SweepParams& P, double2* Zs, const PermView& perm_s,
                                             const Hyper& h, const RngKey& key, const double* z_s,
                                             int pos, int sweep)
{
    const int j = perm_s[pos];
    const double zn = z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0);
    const double pv = 1.0 / (1.0 / h.tau2 + P.Xl2[j] / h.sigma2);
    Zs[pos] = make_double2(sqrt(pv) * zn, pv);
}

// Sliding speculation window: thread t owns the one position of [pos0, pos0+256)
// that is congruent to t modulo 256, so a position keeps its thread (and the cached,
// state-independent part of its decision) when the window restarts after a change.
//
// The decision  u <= kappa(XjTr)  is evaluated through the monotone re-arrangement
//     E := logodds + logdet_j - c1_j * XjTr^2   >=   L := log(u / (1 - u))
// with A_j = logodds + logdet_j, c1_j and L cached per position.  Outside a guard band
// around E == L (and away from the saturation / overflow corners) this is provably the
// same boolean as the reference's floating-point evaluation; inside the band the
// reference formula itself (kappa_of) is evaluated, so every decision is exact.
//
// Deferred column updates.  A change at coordinate j* with increment delta turns q into
// q - delta G[:, j*], but the next decisions only look at q in the positions of the
// window.  Every thread therefore carries q of the position it owns in a register
// (q_pos) and corrects it with ONE gathered Gram entry per change, while the change goes
// on a pending list; after GB_PEND changes (and at the end of a sweep) the pending
// columns are streamed and applied to q in shared memory in a single pass.  A position
// that enters a thread's ownership reads q and replays the pending list.  The fma
// sequence seen by every element of q is the same as with immediate updates, so the
// results are bit-identical to them; what changes is that the latency chain of a change
// is one gather instead of a whole column, and that q is read and written once per
// GB_PEND changes instead of once per change.
template <int DESIGN, bool QSMEM>
__global__ void __maxnreg__(GB_GRAM_REGS) gibbs_gram_kernel(SweepParams P)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SweepShared sh;
    __shared__ GramBcast bc;
    __shared__ int s_first[3];                  // rotating slots: smallest changing offset of a window
    __shared__ int sj[GB_BLOCK];                // coordinate of the position each thread owns
    __shared__ int act_pos[GB_BLOCK];           // positions of the currently active coordinates
    __shared__ int act_n;
    __shared__ int pend_j[GB_PEND];
    __shared__ double pend_d[GB_PEND];
    __shared__ double scratch[40];

    __shared__ int s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = P.p, p2 = P.p2;
    // Persistent CTAs draw (chain, sub-chunk) items from a queue in launch order, so that the
    // SMs stay evenly loaded whatever the number of chains (chains do not come in multiples of
    // the resident CTA count, and one CTA per chain per launch would leave a ragged last wave).
    // A chain's sub-chunks run in order: an item waits for progress[chain] to reach its index.
    // Its predecessor was handed out earlier, i.e. to a CTA that is already running, so the
    // wait always ends.  State crosses CTAs through L2 (ld.cg / fence + flag).
    const int subs = (P.S + P.sub_sweeps - 1) / P.sub_sweeps;
    const int n_items = P.n_chains * subs;
    int parity = 0;
    if (tid < 3) s_first[tid] = GB_BLOCK;
    __syncthreads();
    PH_DECL
  for (;;) {
    if (tid == 0) s_item = atomicAdd(P.queue, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= n_items) break;
    const int c = item % P.n_chains, sub = item / P.n_chains;
    const int s_begin = sub * P.sub_sweeps, s_end = min(P.S, s_begin + P.sub_sweeps);
    if (sub > 0) {
        if (tid == 0) {
            volatile int* pr = P.progress + c;
            while (*pr < sub) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
    }
    double* q_g = P.q + (int64_t)c * p2;
    double* q = QSMEM ? sm : q_g;
    uint32_t* mask_s = (uint32_t*)(sm + (QSMEM ? p2 : 0));

    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    double* beta_g = P.beta + (int64_t)c * p;

    if (QSMEM) for (int k = tid; k < p2; k += GB_BLOCK) q[k] = __ldcg(q_g + k);
    for (int w = tid; w < P.words; w += GB_BLOCK) mask_s[w] = __ldcg(P.mask + (int64_t)c * P.words + w);
    if (tid == 0) {
        sh.sigma2 = __ldcg(P.hstate + GB_HS * c + 0); sh.tau2 = __ldcg(P.hstate + GB_HS * c + 1);
        sh.p0 = __ldcg(P.hstate + GB_HS * c + 2); sh.rr = __ldcg(P.hstate + GB_HS * c + 3);
    }
    __syncthreads();

    long long n_changes = 0, n_windows = 0;     // block-uniform work counters
    for (int s = s_begin; s < s_end; ++s) {
        const int sweep = P.sweep0 + s;
        const int64_t soff = ((int64_t)c * P.S + s) * p;
        // keyed visiting order: computed point-wise below into chain-private scratch; injected: read
        const PermView perm_s = P.perm ? perm_view(P, c, s) : PermView{ P.scrP, (int64_t)c * p, P.perm16 };
        const double* u_s = P.u ? P.u + soff : nullptr;
        const double* z_s = P.z ? P.z + soff : nullptr;
        PH(0);

        // ---- periodic rebuild of q and ||r||^2 from q0 = X^T y (bounds rounding drift)
        if (P.gram_refresh > 0 && sweep > 0 && sweep % P.gram_refresh == 0) {
            for (int k = tid; k < p2; k += GB_BLOCK) q[k] = k < p ? P.q0[k] : 0.0;
            __syncthreads();
            for (int w = 0; w < P.words; ++w) {
                uint32_t m = mask_s[w];
                while (m) {
                    int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    gram_axpy<DESIGN>(P, q, j, __ldcg(beta_g + j));
                }
            }
            __syncthreads();
            double acc = 0.0;                   // ||y - X b||^2 = y'y - sum_j b_j (q0_j + q_j)
            for (int w = tid; w < P.words; w += GB_BLOCK) {
                uint32_t m = mask_s[w];
                while (m) {
                    int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    acc = fma(__ldcg(beta_g + j), P.q0[j] + q[j], acc);
                }
            }
            acc = block_sum(acc, scratch);
            if (tid == 0) sh.rr = P.yy - acc;
        }
        if (tid == 0) sh.logodds = log(sh.p0) - log(1 - sh.p0);
        __syncthreads();
        const Hyper h = { sh.sigma2, sh.tau2, sh.logodds };
        const double ts = h.tau2 / h.sigma2;

        // ---- state-independent per-position quantities of this sweep, computed by the whole
        // block without divergence and parked in (L2-resident) global scratch: the logit of the
        // spike/slab uniform that steers the fast decision, and -- at the positions of the
        // coordinates that are active now, which are exactly the ones that will be active when
        // visited -- the state-independent part of the slab draw (_linear.pyx:130,158).
        float* Ls = P.scrL + (int64_t)c * p;
        double2* Zs = P.scrZ + (int64_t)c * p;
        if (tid == 0) act_n = 0;
        __syncthreads();
        const bool keyed_perm = P.perm == nullptr;
        PermKey pk;
        if (keyed_perm) pk = perm_key(P.k0, P.k1, key.chain, (uint32_t)sweep, (uint32_t)p);
        for (int base = tid; base < p; base += 4 * GB_BLOCK) {
            int jj[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pos = base + i * GB_BLOCK;
                if (pos >= p) jj[i] = -1;
                else if (keyed_perm) {
                    jj[i] = (int)perm_at(pk, (uint32_t)pos, (uint32_t)p);
                    if (P.perm16) reinterpret_cast<uint16_t*>(P.scrP)[(int64_t)c * p + pos] = (uint16_t)jj[i];
                    else reinterpret_cast<int32_t*>(P.scrP)[(int64_t)c * p + pos] = jj[i];
                } else jj[i] = perm_s[pos];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pos = base + i * GB_BLOCK;
                if (pos < p) {
                    const double u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                    const bool ex = !(u > 1e-4 && u < 1.0 - 1e-4);
                    Ls[pos] = ex ? __int_as_float(0x7fc00000) : __logf((float)u) - __logf((float)(1.0 - u));
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = jj[i];
                if (j >= 0 && ((mask_s[j >> 5] >> (j & 31)) & 1u)) {
                    const int slot = atomicAdd(&act_n, 1);
                    if (slot < GB_BLOCK) act_pos[slot] = base + i * GB_BLOCK;
                    else slab_prepare(P, Zs, perm_s, h, key, z_s, base + i * GB_BLOCK, sweep);   // list full (early burn-in)
                }
            }
        }
        __syncthreads();
        if (tid < min(act_n, GB_BLOCK)) slab_prepare(P, Zs, perm_s, h, key, z_s, act_pos[tid], sweep);
        __syncthreads();

        // per-position cache: filled when a position enters a thread's ownership, i.e. one
        // window ahead of its first evaluation.  f_A, f_c1, f_L only steer the guarded fast
        // decision, so single precision (and the fast intrinsics) is enough for them: the
        // guard band below is 1e-4 relative.  Anything that reaches a coefficient or an
        // exact decision stays in fp64.
        int c_pos = -1, c_j = 0, n_j = 0, nn_j = 0, npend = 0;
        bool c_act = false;
        int n_def = 0;                    // pending changes still to be replayed into q_pos
        double d_g[GB_PEND], d_d[GB_PEND];
        float f_A = 0.f, f_c1 = 0.f, f_L = 0.f, n_L = 0.f;
        double c_xl2 = 0.0, n_xl2 = 0.0, c_bold = 0.0;
        double2 c_zp = make_double2(0.0, 0.0);   // (sqrt(post_var) * z, post_var) of an active position
        double q_pos = 0.0;               // q[c_j] with every change so far applied
        const float f_logodds = (float)h.logodds, f_tau2 = (float)h.tau2, f_sigma2 = (float)h.sigma2;
        const float f_ts = (float)ts;
        // needs n_j, n_xl2, n_L = perm / Xl2 / logit of `pos` (prefetched).  If jnow >= 0 a change
        // at coordinate jnow is being decided: its Gram entry for the new position is fetched too.
        auto fill = [&](int pos, int jnow, bool publish = true) -> double {
            c_pos = pos;
            c_j = n_j;
            if (publish) sj[tid] = n_j;   // (the owner of a change publishes after the barrier, see below)
            const double gnow = jnow >= 0 ? gram_entry<DESIGN>(P, jnow, c_j) : 0.0;
            // the pending changes are not in q yet: request their Gram entries now, apply them
            // (replay) right before q_pos is next used, so that the loads are never waited for here
#pragma unroll
            for (int k = 0; k < GB_PEND; ++k)
                if (k < npend) { d_g[k] = gram_entry<DESIGN>(P, pend_j[k], c_j); d_d[k] = pend_d[k]; }
            n_def = npend;
            q_pos = q[c_j];
            c_xl2 = n_xl2;
            f_L = n_L;
            c_act = (mask_s[c_j >> 5] >> (c_j & 31)) & 1u;
            if (c_act) { c_bold = __ldcg(beta_g + c_j); c_zp = __ldcg(Zs + pos); } else c_bold = 0.0;
            // two-deep prefetch of the visiting order (and one-deep of what hangs off it), so that
            // no dependent load is ever waited for: nn_j belongs to pos + 2*256
            n_j = nn_j;
            if (pos + GB_BLOCK < p) { n_xl2 = P.Xl2[n_j]; n_L = __ldcg(Ls + pos + GB_BLOCK); }
            if (pos + 2 * GB_BLOCK < p) nn_j = perm_s[pos + 2 * GB_BLOCK];
            const float xl2f = (float)c_xl2;
            f_A = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
            f_c1 = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
            return gnow;
        };
        auto replay = [&]() {             // apply the deferred pending changes, oldest first
#pragma unroll
            for (int k = 0; k < GB_PEND; ++k) if (k < n_def) q_pos = fma(-d_d[k], d_g[k], q_pos);
            n_def = 0;
        };
        if (tid < p) {
            n_j = perm_s[tid]; n_xl2 = P.Xl2[n_j]; n_L = __ldcg(Ls + tid);
            if (tid + GB_BLOCK < p) nn_j = perm_s[tid + GB_BLOCK];
            fill(tid, -1);
        }
        PH(1);

        int pos0 = 0;
        while (pos0 < p) {
            ++n_windows;
            const int off = (GB_BLOCK & (GB_BLOCK - 1)) == 0 ? ((tid - pos0) & (GB_BLOCK - 1)) : gb_mod(tid - pos0 % GB_BLOCK);
            const int pos = pos0 + off;
            bool change = false, spike = true;
            double XjTr = 0.0;
            if (pos < p) {
                if (pos != c_pos) {                           // safety net; not reached by construction
                    n_j = perm_s[pos]; n_xl2 = P.Xl2[n_j]; n_L = __ldcg(Ls + pos);
                    if (pos + GB_BLOCK < p) nn_j = perm_s[pos + GB_BLOCK];
                    fill(pos, -1);
                }
                replay();
                XjTr = c_act ? q_pos + c_bold * c_xl2 : q_pos;      // X_j . (r + beta_j X_j)
                const double x2c = (double)f_c1 * (XjTr * XjTr);
                const double E = (double)f_A - x2c;
                const double margin = E - (double)f_L;
                const double guard = 1e-4 * (1.0 + fabs((double)f_A) + x2c + fabs((double)f_L));
                // (a NaN f_L -- uniform outside [1e-4, 1-1e-4] -- fails the first test)
                if (!(fabs(margin) > guard) || !(fmax(E, (double)f_L) < 20.0)) {
                    const double c_u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                    spike = (c_u <= kappa_of(XjTr, c_xl2, h));   // the reference's own formula
                } else {
                    spike = margin > 0.0;
                }
                change = c_act || !spike;
            }
            // first changed position in window order (= smallest cyclic offset from pos0): every
            // warp posts the smallest offset among its lanes with one shared-memory atomicMin
            const unsigned m = __ballot_sync(0xffffffffu, change);
            const int r = gb_mod(pos0);
            if (m && lane == 0) {
                int l;
                if (warp == (r >> 5)) {                 // the warp that holds the window start: lanes
                    const unsigned hi = m & (0xffffffffu << (r & 31));   // at/after it come first
                    l = hi ? __ffs(hi) - 1 : __ffs(m) - 1;
                } else l = __ffs(m) - 1;
                atomicMin(&s_first[parity], gb_mod(32 * warp + l - r));
            }
            PH(2);
            __syncthreads();
            PH(3);
            int first = s_first[parity];
            if (first >= GB_BLOCK) first = -1;
            if (tid == 0) s_first[parity == 0 ? 2 : parity - 1] = GB_BLOCK;   // slot of the window after next
            parity = parity == 2 ? 0 : parity + 1;
            PH(4);
            if (first < 0) {                                  // the whole window was null
                pos0 += GB_BLOCK;
                if (pos + GB_BLOCK < p) fill(pos + GB_BLOCK, -1); else c_pos = -1;
                PH(5);
                continue;
            }
            ++n_changes;
            // Everybody knows the changing coordinate without waiting for its owner: the Gram
            // entry that corrects this thread's own q_pos is requested before the slab draw and
            // before the leaving positions are re-filled, so its latency hides behind that math.
            const int jstar = sj[gb_mod(pos0 + first)];
            double gj = 0.0;
            PH(6);
            if (off > first) {
                if (c_pos >= 0) gj = gram_entry<DESIGN>(P, jstar, c_j);
            } else if (off == first) {                        // the owner of the changing position
                double bnew = 0.0;
                if (!spike) {
                    if (c_act) bnew = c_zp.x + c_zp.y * XjTr / h.sigma2;  // = slab_draw(...)
                    else bnew = slab_draw(XjTr, c_xl2, h, z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0));
                }
                const double delta = bnew - c_bold;
                bc.delta = delta;
                pend_j[npend] = c_j; pend_d[npend] = delta;
                // ||r - delta X_j||^2 = ||r||^2 - 2 delta X_j.r + delta^2 ||X_j||^2
                sh.rr = sh.rr - 2.0 * delta * q_pos + delta * delta * c_xl2;
                __stcg(beta_g + c_j, bnew);
                if (bnew != 0) mask_s[c_j >> 5] |= (1u << (c_j & 31));
                else mask_s[c_j >> 5] &= ~(1u << (c_j & 31));
                // a column that is not L2-resident is on its way by the time it is flushed
                if (DESIGN == BLIPGPU_DESIGN_DENSE)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.G + (int64_t)c_j * P.ldg),
                                 "r"((unsigned)(p2 * sizeof(double))) : "memory");
            }
            // positions up to and including the changed one leave the window: their threads take
            // over the positions one window further (the owner after its draw, in the same
            // divergent pass as the other lanes of its warp).  The current change is not on the
            // pending list they replay yet; it reaches them through gj below like everybody else.
            // The owner's slot of sj is the one everybody reads jstar from: it is rewritten only after
            // the barrier, when all those reads are done.
            if (off <= first) {
                if (pos + GB_BLOCK < p) gj = fill(pos + GB_BLOCK, jstar, off != first); else c_pos = -1;
            }
            PH(7);
            __syncthreads();
            PH(8);
            if (off == first && c_pos >= 0) sj[tid] = c_j;
            replay();
            q_pos = fma(-bc.delta, gj, q_pos);
            ++npend;
            PH(9);
            if (npend == GB_PEND) {
                gram_flush<DESIGN>(P, q, pend_j, pend_d, npend);
                npend = 0;
                __syncthreads();
            }
            PH(10);
            pos0 += first + 1;
            PH(11);
        }
        if (npend > 0) {                                      // block-uniform
            __syncthreads();
            gram_flush<DESIGN>(P, q, pend_j, pend_d, npend);
        }
        __syncthreads();

        sweep_epilogue(P, &sh, mask_s, beta_g, scratch, key, c, s, sh.rr, nullptr, false);
        __syncthreads();
    }
    if (QSMEM) for (int k = tid; k < p2; k += GB_BLOCK) q_g[k] = q[k];
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2;
        P.hstate[GB_HS * c + 2] = sh.p0; P.hstate[GB_HS * c + 3] = sh.rr;
        P.hstate[GB_HS * c + 4] = __ldcg(P.hstate + GB_HS * c + 4) + (double)n_changes;
        P.hstate[GB_HS * c + 5] = __ldcg(P.hstate + GB_HS * c + 5) + (double)n_windows;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) { volatile int* pr = P.progress + c; *pr = sub + 1; }
  }
#ifdef BLIP_PHASE_CLOCKS
    if (tid == 0 && blockIdx.x == 0) {
        printf("cta 0 | epi+sweepstart %lld fill0 %lld | eval %lld sync1 %lld find %lld nullfill %lld | jstar %lld owner/fill/gather %lld sync2 %lld apply %lld flush %lld ownerfill %lld\n",
               ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[6], ph_t[7], ph_t[8], ph_t[9], ph_t[10], ph_t[11]);
    }
#endif
}
