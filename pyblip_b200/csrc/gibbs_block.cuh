// gibbs_block.cuh -- device side of the BLOCK spike-and-slab Gibbs sampler (bsize > 1), sm_100a.
//
// Semantics follow /root/reference/pyblip/linear/_linear_multi.pyx:271-932: the coordinates are
// partitioned into blocks of `bsize` consecutive (wrapping) indices, visited in an order fixed
// by the shuffle of sweep 0; a visit zeroes the block, scores every non-empty sub-model of at
// most `max_signals_per_block` members (:460-489), draws one (:531-533) and, if it is not the
// empty model, draws its coefficients from the conditional Gaussian (:554-834).
//
// Exact speculation carries over from the single-site kernels: a block that holds no non-zero
// coefficient and whose draw is the empty model leaves the chain state untouched, so a window
// of upcoming blocks is scored in parallel against the current state -- one lane per sub-model,
// 32 / npad blocks per warp (npad = models per block rounded up to a power of two) -- the first
// block that changes anything is applied by the whole CTA, and the window restarts right after
// it.  In the gram form a block that holds coefficients is scored in the window as well, on the
// statistics of the zeroed block, and the apply step reuses those model probabilities.
//
// Two formulations (template GRAM):
//   residual  state r = y - X beta in shared memory; X_A^T r costs bsize dots of length n per
//             block; required for probit (a non-empty draw redraws all n latents, :875-884);
//   gram      state q = X^T r in shared memory; X_A^T r is a lookup and a coefficient change
//             streams one Gram column, like gibbs_gram_kernel.
#pragma once
#include "gibbs_kernels.cuh"

constexpr int BK_MAXB = 8;           // largest block size (2^8 - 1 sub-models per block)

// log P(model) up to the common constant, _linear_multi.pyx:463-489.  A: bsize x bsize block of
// X^T X, x: X_A^T r.  Returns NaN if the Cholesky factorisation of I + tau2/sigma2 X_A^T X_A fails
// (the reference raises RuntimeError, :234-236).
__device__ __forceinline__ double block_model_lp(uint32_t mask, int bsize, const double* A, const double* x,
                                                 double tau2, double sigma2, double logp0, double log1p0)
{
    int mem[BK_MAXB];
    int m = 0;
    for (int jj = 0; jj < bsize; ++jj) if (mask & (1u << jj)) mem[m++] = jj;
    const double cm = tau2 / sigma2;
    double Q[BK_MAXB * BK_MAXB], yv[BK_MAXB];
    for (int a = 0; a < m; ++a)
        for (int b = 0; b <= a; ++b) Q[a * BK_MAXB + b] = (a == b ? 1.0 : 0.0) + cm * A[mem[a] * bsize + mem[b]];
    double logdet = 0.0;
    for (int j = 0; j < m; ++j) {                       // lower Cholesky (dpotrf 'L')
        double d = Q[j * BK_MAXB + j];
        for (int k = 0; k < j; ++k) d -= Q[j * BK_MAXB + k] * Q[j * BK_MAXB + k];
        if (!(d > 0.0)) return nan("");
        d = sqrt(d);
        Q[j * BK_MAXB + j] = d;
        logdet += 2.0 * log(d);
        for (int i = j + 1; i < m; ++i) {
            double s = Q[i * BK_MAXB + j];
            for (int k = 0; k < j; ++k) s -= Q[i * BK_MAXB + k] * Q[j * BK_MAXB + k];
            Q[i * BK_MAXB + j] = s / d;
        }
    }
    double nrm2 = 0.0;                                  // || L^-1 X_A^T r ||^2 by forward substitution
    for (int a = 0; a < m; ++a) {
        double s = x[mem[a]];
        for (int k = 0; k < a; ++k) s -= Q[a * BK_MAXB + k] * yv[k];
        yv[a] = s / Q[a * BK_MAXB + a];
        nrm2 += yv[a] * yv[a];
    }
    double nrm = sqrt(nrm2);                            // dnrm2, then squared (:479-480)
    double lp = tau2 * nrm * nrm / (2 * sigma2 * sigma2) - logdet / 2;
    lp += (bsize - m) * logp0;
    lp += m * log1p0;
    return lp;
}

// The same score with the block size as a compile-time constant: the model is embedded in the
// full BS x BS matrix with identity rows for the members it leaves out.  Every operation on a
// member entry is the one block_model_lp performs (the extra terms are exact zeros), so the
// result is bit-identical, but all loops unroll and the small matrices live in registers instead
// of dynamically indexed local memory.
template <int BS>
__device__ __forceinline__ double block_model_lp_t(uint32_t mask, const double* A, const double* x,
                                                   double tau2, double sigma2, double logp0, double log1p0)
{
    const double cm = tau2 / sigma2;
    double Q[BS][BS], yv[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            const bool in = ((mask >> a) & 1u) && ((mask >> b) & 1u);
            Q[a][b] = (a == b ? 1.0 : 0.0) + (in ? cm * A[a * BS + b] : 0.0);
        }
    double logdet = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < BS; ++j) {
        double d = Q[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= Q[j][k] * Q[j][k];
        if (!(d > 0.0)) ok = false;
        d = sqrt(d);
        Q[j][j] = d;
        if ((mask >> j) & 1u) logdet += 2.0 * log(d);
#pragma unroll
        for (int i = j + 1; i < BS; ++i) {
            double sacc = Q[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) sacc -= Q[i][k] * Q[j][k];
            // an entry outside the model is an exact 0 / d; dividing 1.0 instead and discarding the
            // quotient keeps the whole warp out of the division's zero-numerator slow path
            const bool in = ((mask >> i) & 1u) && ((mask >> j) & 1u);
            double num = in ? sacc : 1.0;
            asm volatile("" : "+d"(num));          // keeps the compiler from dividing the zero after all
            const double qt = num / d;
            Q[i][j] = in ? qt : 0.0;
        }
    }
    if (!ok) return nan("");
    double nrm2 = 0.0;
#pragma unroll
    for (int a = 0; a < BS; ++a) {
        double sacc = ((mask >> a) & 1u) ? x[a] : 0.0;
#pragma unroll
        for (int k = 0; k < a; ++k) sacc -= Q[a][k] * yv[k];
        const bool in = (mask >> a) & 1u;
        double num = in ? sacc : 1.0;
        asm volatile("" : "+d"(num));
        const double qt = num / Q[a][a];
        yv[a] = in ? qt : 0.0;
        if (in) nrm2 += yv[a] * yv[a];
    }
    const int m = __popc(mask);
    double nrm = sqrt(nrm2);
    double lp = tau2 * nrm * nrm / (2 * sigma2 * sigma2) - logdet / 2;
    lp += (BS - m) * logp0;
    lp += m * log1p0;
    return lp;
}

__device__ __noinline__ double block_model_lp_any(uint32_t mask, int bsize, const double* A, const double* x,
                                                  double tau2, double sigma2, double logp0, double log1p0)
{
    switch (bsize) {
        case 2: return block_model_lp_t<2>(mask, A, x, tau2, sigma2, logp0, log1p0);
        case 3: return block_model_lp_t<3>(mask, A, x, tau2, sigma2, logp0, log1p0);
        case 4: return block_model_lp_t<4>(mask, A, x, tau2, sigma2, logp0, log1p0);
        case 5: return block_model_lp_t<5>(mask, A, x, tau2, sigma2, logp0, log1p0);
        case 6: return block_model_lp_t<6>(mask, A, x, tau2, sigma2, logp0, log1p0);
        default: return block_model_lp(mask, bsize, A, x, tau2, sigma2, logp0, log1p0);
    }
}

// Scores all models of one block with one warp: probs[mj] = exp(lp_mj - max(0, max lp)) in shared
// memory (model 0 = empty), returns the sequentially accumulated normaliser (same order as
// _weighted_choice, :55-79) in every lane; *bad is set if a factorisation failed.
__device__ __forceinline__ double block_score(const SweepParams& P, const double* A, const double* x, double tau2,
                                              double sigma2, double logp0, double log1p0, double* probs, bool* bad)
{
    const int lane = threadIdx.x & 31;
    double mx = 0.0;                                    // max_log_prob starts at 0.0 (:62)
    bool fail = false;
    for (int mj = lane; mj < P.nmodel; mj += 32) {
        double lp = mj == 0 ? P.bsize * logp0
                            : block_model_lp_any(P.model_masks[mj - 1], P.bsize, A, x, tau2, sigma2, logp0, log1p0);
        if (lp != lp) fail = true;
        probs[mj] = lp;
        mx = fmax(lp, mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    for (int mj = lane; mj < P.nmodel; mj += 32) probs[mj] = exp(probs[mj] - mx);
    __syncwarp();
    double denom = 0.0;
    if (lane == 0) for (int mj = 0; mj < P.nmodel; ++mj) denom += probs[mj];
    denom = __shfl_sync(0xffffffffu, denom, 0);
    *bad = __any_sync(0xffffffffu, fail);
    return denom;
}

// Conditional Gaussian draw of the selected model (_linear_multi.pyx:535-834), one thread.
// out[k] = new coefficient of block member k (0 for members outside the model); returns false
// if a Cholesky factorisation fails.
__device__ inline bool block_draw(uint32_t mask, int bsize, const double* A, const double* x, double tau2,
                                  double sigma2, const double* zn, double* out)
{
    int mem[BK_MAXB];
    int m = 0;
    for (int jj = 0; jj < bsize; ++jj) { out[jj] = 0.0; if (mask & (1u << jj)) mem[m++] = jj; }
    const double cm = tau2 / sigma2;
    if (m == 1) {
        // one member (the usual draw in the stationary regime): the operations of the general path
        // below on 1 x 1 matrices, in the same order, without the local-memory matrices
        const double a11 = A[mem[0] * bsize + mem[0]], x1 = x[mem[0]];
        double l = 1.0 + cm * a11;
        if (!(l > 0.0)) return false;
        l = sqrt(l);
        double mu1 = x1 / l; mu1 = mu1 / l;
        double s1 = 0.0; s1 += a11 * mu1;
        const double t1 = cm * s1;
        mu1 = t1 - x1; mu1 *= -1 * tau2 / sigma2;
        double w1 = a11 / l; w1 = w1 / l;
        const double c2s = -1 * tau2 * tau2 / sigma2, c3s = tau2 * tau2 * tau2 / (sigma2 * sigma2);
        double vs = 0.0; vs += a11 * w1;
        double v = tau2; v += c2s * a11; v += c3s * vs;
        if (!(v > 0.0)) return false;
        v = sqrt(v);
        double o = 0.0; o += v * zn[0];
        out[mem[0]] = o + mu1;
        return true;
    }
    double Am[BK_MAXB * BK_MAXB], L[BK_MAXB * BK_MAXB], W[BK_MAXB * BK_MAXB], V[BK_MAXB * BK_MAXB];
    double xm[BK_MAXB], mu[BK_MAXB], t[BK_MAXB];
    for (int a = 0; a < m; ++a) {
        xm[a] = x[mem[a]];
        for (int b = 0; b < m; ++b) {
            Am[a * BK_MAXB + b] = A[mem[a] * bsize + mem[b]];
            L[a * BK_MAXB + b] = (a == b ? 1.0 : 0.0) + cm * Am[a * BK_MAXB + b];
        }
    }
    auto chol = [&](double* M) -> bool {
        for (int j = 0; j < m; ++j) {
            double d = M[j * BK_MAXB + j];
            for (int k = 0; k < j; ++k) d -= M[j * BK_MAXB + k] * M[j * BK_MAXB + k];
            if (!(d > 0.0)) return false;
            d = sqrt(d);
            M[j * BK_MAXB + j] = d;
            for (int i = j + 1; i < m; ++i) {
                double s = M[i * BK_MAXB + j];
                for (int k = 0; k < j; ++k) s -= M[i * BK_MAXB + k] * M[j * BK_MAXB + k];
                M[i * BK_MAXB + j] = s / d;
            }
        }
        return true;
    };
    auto solve = [&](const double* rhs, double* sol) {   // (L L^T) sol = rhs
        for (int a = 0; a < m; ++a) {
            double s = rhs[a];
            for (int k = 0; k < a; ++k) s -= L[a * BK_MAXB + k] * sol[k];
            sol[a] = s / L[a * BK_MAXB + a];
        }
        for (int a = m - 1; a >= 0; --a) {
            double s = sol[a];
            for (int k = a + 1; k < m; ++k) s -= L[k * BK_MAXB + a] * sol[k];
            sol[a] = s / L[a * BK_MAXB + a];
        }
    };
    if (!chol(L)) return false;
    solve(xm, mu);                                      // Q_A^-1 X_A^T r            (:554-567)
    for (int a = 0; a < m; ++a) {                       // tau2/sigma2 (x - cm A mu) (:571-597)
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += Am[a * BK_MAXB + k] * mu[k];
        t[a] = cm * s;
    }
    for (int a = 0; a < m; ++a) { mu[a] = t[a] - xm[a]; mu[a] *= -1 * tau2 / sigma2; }
    for (int b = 0; b < m; ++b) {                       // W = Q_A^-1 A, column by column (:599-640)
        double rhs[BK_MAXB], sol[BK_MAXB];
        for (int a = 0; a < m; ++a) rhs[a] = Am[a * BK_MAXB + b];
        solve(rhs, sol);
        for (int a = 0; a < m; ++a) W[a * BK_MAXB + b] = sol[a];
    }
    const double c2 = -1 * tau2 * tau2 / sigma2, c3 = tau2 * tau2 * tau2 / (sigma2 * sigma2);
    for (int a = 0; a < m; ++a)                         // V = tau2 I - tau2^2/sigma2 A + tau2^3/sigma2^2 A W (:641-754)
        for (int b = 0; b <= a; ++b) {
            double vs = 0.0;
            for (int k = 0; k < m; ++k) vs += Am[a * BK_MAXB + k] * W[k * BK_MAXB + b];
            double v = (a == b) ? tau2 : 0.0;
            v += c2 * Am[a * BK_MAXB + b];
            v += c3 * vs;
            V[a * BK_MAXB + b] = v;
        }
    if (!chol(V)) return false;                         // (:755-766)
    for (int a = 0; a < m; ++a) {                       // beta = L_V z + mu (:782-834)
        double s = 0.0;
        for (int k = 0; k <= a; ++k) s += V[a * BK_MAXB + k] * zn[k];
        out[mem[a]] = s + mu[a];
    }
    return true;
}

struct BlockShared {
    int first, bi, mj;
    int mem[BK_MAXB];
    double bold[BK_MAXB], bnew[BK_MAXB], xatr[BK_MAXB], A[BK_MAXB * BK_MAXB];
    unsigned flags[GB_NWARP];
    unsigned failb[GB_NWARP];          // lanes of each warp whose window score hit a failed factorisation
    double zn[BK_MAXB];                // normals of the coefficient draw
    double u[GB_NWARP][8];             // choice uniforms of the blocks of the window (drawn while their loads fly)
    double cq[32];                     // model probabilities of the applied block, normalised
};

#ifdef BK_PHASE_CLOCKS
#define BPH_DECL long long bph[10] = {0,0,0,0,0,0,0,0,0,0}; long long bph_last = clock64();
#define BPH(i) do { if (threadIdx.x == 0) { long long _n = clock64(); bph[i] += _n - bph_last; bph_last = _n; } } while (0)
#define BPH_PRINT do { if (threadIdx.x == 0 && blockIdx.x == 0 && (sweep % 16 == 0)) \
    printf("sweep %d windows %lld applied %lld | start %lld score %lld find %lld | stats %lld rescore+draw %lld writeback %lld | epilogue %lld\n", sweep, \
           n_windows, n_changes, bph[0], bph[1], bph[2], bph[3], bph[4], bph[5], bph[6]); } while (0)
#else
#define BPH_DECL
#define BPH(i) do {} while (0)
#define BPH_PRINT do {} while (0)
#endif

template <bool GRAM, bool PROBIT, int DESIGN = BLIPGPU_DESIGN_DENSE>
__global__ void __launch_bounds__(GB_BLOCK, 2) gibbs_block_kernel(SweepParams P)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SweepShared sh;
    __shared__ BlockShared bs;
    __shared__ double scratch[40];
    __shared__ double w_x[GB_NWARP][16];         // X_A^T r of the (up to 8) blocks a warp scores
    __shared__ double w_A[GB_NWARP][BK_MAXB * BK_MAXB];

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.n, p = P.p, n2 = P.n2, p2 = P.p2, bsize = P.bsize, nblock = P.nblock;
    // layout: state (q | r [mu y]) | per-warp model probabilities | mask
    double* st_s = sm;
    const int state_len = GRAM ? p2 : (PROBIT ? 3 : 1) * n2;
    double* r_s = GRAM ? nullptr : st_s;
    double* mu_s = PROBIT ? st_s + n2 : nullptr;
    double* y_s = PROBIT ? st_s + 2 * n2 : nullptr;
    double* q = GRAM ? st_s : nullptr;
    double* probs_all = st_s + state_len;                       // [GB_NWARP][nmodel_pad]
    const int nmodel_pad = max(32, (P.nmodel + 1) & ~1);
    double* probs = probs_all + warp * nmodel_pad;
    uint32_t* mask_s = (uint32_t*)(probs_all + GB_NWARP * nmodel_pad);

    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    double* beta_g = P.beta + (int64_t)c * p;
    const int32_t* blocks = P.blocks + (int64_t)c * nblock * bsize;

    if (GRAM) for (int k = tid; k < p2; k += GB_BLOCK) q[k] = P.q[(int64_t)c * p2 + k];
    else for (int k = tid; k < n2; k += GB_BLOCK) {
        r_s[k] = P.r[(int64_t)c * n2 + k];
        if (PROBIT) { mu_s[k] = P.mu[(int64_t)c * n2 + k]; y_s[k] = P.ylat[(int64_t)c * n2 + k]; }
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) mask_s[w] = P.mask[(int64_t)c * P.words + w];
    if (tid == 0) {
        sh.sigma2 = P.hstate[GB_HS * c + 0]; sh.tau2 = P.hstate[GB_HS * c + 1];
        sh.p0 = P.hstate[GB_HS * c + 2]; sh.rr = P.hstate[GB_HS * c + 3];
    }
    long long n_changes = 0, n_windows = 0;
    uint32_t event = PROBIT ? P.events[2 * c] : 0u;
    bool dirty = PROBIT ? (P.events[2 * c + 1] != 0u) : false;
    __syncthreads();

    // X_j . r for the whole warp (residual form)
    auto warp_dot = [&](int j) -> double {
        const double* xcol = P.XT + (int64_t)j * P.ldx;
        double acc = 0.0;
        for (int k = 2 * lane; k < n2; k += 64) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(xcol + k));
            const double2 rk = *reinterpret_cast<const double2*>(r_s + k);
            acc = fma(x.x, rk.x, acc); acc = fma(x.y, rk.y, acc);
        }
        return warp_sum(acc);
    };
    // state += coef * column j (residual: r += coef X_j, probit also mu -= coef X_j; gram: q += coef G[:,j])
    auto add_column = [&](int j, double coef) {
        if (GRAM) {
            if (tid == 0) sh.rr = sh.rr + 2.0 * coef * q[j] + coef * coef * P.Xl2[j];   // ||r + coef X_j||^2
            __syncthreads();
            gram_axpy<DESIGN>(P, q, j, -coef);
        } else {
            const double* xcol = P.XT + (int64_t)j * P.ldx;
            for (int k = tid; k < n; k += GB_BLOCK) {
                const double x = __ldg(xcol + k);
                r_s[k] = fma(coef, x, r_s[k]);
                if (PROBIT) mu_s[k] = __dadd_rn(__dmul_rn(-coef, x), mu_s[k]);   // rounded product: exact cancellation
                                                                                 // to 0.0 (see gibbs_residual_kernel)
            }
        }
        __syncthreads();
    };

    for (int s = 0; s < P.S; ++s) {
        const int sweep = P.sweep0 + s;
        BPH_DECL
        const double* u_s = P.u ? P.u + ((int64_t)c * P.S + s) * nblock : nullptr;
        const double* z_s = P.z ? P.z + ((int64_t)c * P.S + s) * nblock * bsize : nullptr;
        if (GRAM && P.gram_refresh > 0 && sweep > 0 && sweep % P.gram_refresh == 0) {
            // periodic rebuild of q and ||r||^2 from q0 = X^T y (bounds rounding drift)
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
            double acc = 0.0;
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
        const double tau2 = sh.tau2, sigma2 = sh.sigma2;
        const double logp0 = log(sh.p0), log1p0 = log(1 - sh.p0);       // :339-340, :912-913

        BPH(0);
        int pos0 = 0;
        while (pos0 < nblock) {
            ++n_windows;
            // ---- speculative scoring.  A block needs one lane per model, so a warp scores
            // G = 32 / npad blocks at once (npad = models per block rounded up to a power of two;
            // bsize 3 -> 8 models -> 4 blocks per warp); beyond 32 models a warp takes one block
            // and its lanes stride over the models.
            const int npad = P.npad;
            const int G = P.blocks_per_warp;                            // <= 32 / npad, bounded by the scratch
            const int WB = GB_NWARP * G;                                // blocks per window
            unsigned changed = 0u;                                      // bit g: block g of this warp changes
            if (npad <= 32) {
                const int grp = lane / npad, gl = lane - grp * npad;
                const int bi_g = pos0 + warp * G + grp;
                const bool valid = grp < G && bi_g < nblock;
                const unsigned gmask = (npad == 32 ? 0xffffffffu : ((1u << npad) - 1u)) << (grp * npad);
                const int j = (valid && gl < bsize) ? blocks[bi_g * bsize + gl] : 0;
                // the choice uniform does not depend on the state: drawn here, under the loads
                double u_blk = 0.0;
                if (valid && gl == 0) {
                    u_blk = u_s ? u_s[bi_g] : rng_uniform(key, KIND_BLOCK, bi_g, sweep, 0);
                    bs.u[warp][grp] = u_blk;
                }
                const bool act = valid && gl < bsize && ((mask_s[j >> 5] >> (j & 31)) & 1u);
                const unsigned actb = __ballot_sync(0xffffffffu, act);
                const bool certain = (actb & gmask) != 0u;             // must be zeroed: a certain change
                // gram form: a block that holds coefficients is scored as well, on the statistics of
                // the zeroed block (q_m + sum_k beta_k G[m][k], the chain of the apply step), so that
                // the apply step below never scores again
                const bool score = valid && (GRAM || !certain);
                const unsigned scoreb = __ballot_sync(0xffffffffu, score && gl == 0);
                double* gx = w_x[warp] + grp * bsize;
                double* gA = w_A[warp] + grp * bsize * bsize;
                const int j0 = __shfl_sync(0xffffffffu, j, grp * npad);
                const double* gb = P.gb + (int64_t)(j0 / bsize) * bsize * bsize;
                double xg = 0.0;
                if (GRAM) {
                    const double bold = act ? __ldcg(beta_g + j) : 0.0;
                    if (score && gl < bsize) xg = q[j];
                    if (__any_sync(0xffffffffu, certain && valid)) {
                        for (int k = 0; k < bsize; ++k) {
                            const double bk = __shfl_sync(0xffffffffu, bold, grp * npad + k);
                            if (score && gl < bsize && bk != 0) xg = fma(bk, gb[gl * bsize + k], xg);
                        }
                    }
                }
                if (score) {
                    for (int e = gl; e < bsize * bsize; e += npad) gA[e] = gb[e];
                    if (GRAM && gl < bsize) gx[gl] = xg;
                }
                if (!GRAM) {                                            // dots need the whole warp
                    for (int g = 0; g < G; ++g) {
                        if (!((scoreb >> (g * npad)) & 1u)) continue;   // warp-uniform
                        for (int k = 0; k < bsize; ++k) {
                            const int jk = __shfl_sync(0xffffffffu, j, g * npad + k);
                            const double d = warp_dot(jk);
                            if (lane == 0) w_x[warp][g * bsize + k] = d;
                        }
                    }
                }
                __syncwarp();
                double lp = 0.0;
                bool fail = false;
                if (score && gl < P.nmodel) {
                    lp = gl == 0 ? bsize * logp0
                                 : block_model_lp_any(P.model_masks[gl - 1], bsize, gA, gx, tau2, sigma2, logp0, log1p0);
                    fail = lp != lp;
                }
                double mx = fmax(lp, 0.0);                              // max_log_prob starts at 0.0 (:62)
                for (int o = npad >> 1; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                double* gp = probs + grp * npad;                        // 32 doubles per warp
                if (score && gl < P.nmodel) gp[gl] = exp(lp - mx);
                const unsigned failb = __ballot_sync(0xffffffffu, fail);
                if (lane == 0) bs.failb[warp] = failb;
                __syncwarp();
                bool chg = false;
                if (valid && gl == 0) {
                    if (certain || (failb & gmask)) chg = true;
                    else {
                        double denom = 0.0;
                        for (int mj = 0; mj < P.nmodel; ++mj) denom += gp[mj];
                        chg = !(0.0 + gp[0] / denom >= u_blk);          // model 0 is drawn iff its share reaches u
                    }
                }
                const unsigned chgb = __ballot_sync(0xffffffffu, chg);
                for (int g = 0; g < G; ++g) if ((chgb >> (g * npad)) & 1u) changed |= 1u << g;
            } else {
                const int bi_w = pos0 + warp;
                bool change = false, bad = false;
                if (bi_w < nblock) {
                    const int j = lane < bsize ? blocks[bi_w * bsize + lane] : 0;
                    const bool act = lane < bsize && ((mask_s[j >> 5] >> (j & 31)) & 1u);
                    if (__any_sync(0xffffffffu, act)) change = true;
                    else {
                        const int j0 = __shfl_sync(0xffffffffu, j, 0);
                        const double* gb = P.gb + (int64_t)(j0 / bsize) * bsize * bsize;
                        for (int e = lane; e < bsize * bsize; e += 32) w_A[warp][e] = gb[e];
                        for (int k = 0; k < bsize; ++k) {
                            const int jk = __shfl_sync(0xffffffffu, j, k);
                            const double d = GRAM ? q[jk] : warp_dot(jk);
                            if (lane == 0) w_x[warp][k] = d;
                        }
                        __syncwarp();
                        const double denom = block_score(P, w_A[warp], w_x[warp], tau2, sigma2, logp0, log1p0, probs, &bad);
                        const double u = u_s ? u_s[bi_w] : rng_uniform(key, KIND_BLOCK, bi_w, sweep, 0);
                        change = bad || !(0.0 + probs[0] / denom >= u);
                    }
                }
                changed = change ? 1u : 0u;
            }
            BPH(1);
            if (lane == 0) bs.flags[warp] = changed;
            __syncthreads();
            int first = -1;
#pragma unroll
            for (int w = GB_NWARP - 1; w >= 0; --w) if (bs.flags[w]) first = w * G + __ffs(bs.flags[w]) - 1;
            __syncthreads();
            BPH(2);
            if (first < 0) { pos0 += WB; continue; }

            // ---- apply block bi exactly like the reference
            ++n_changes;
            const int bi = pos0 + first;
            if (tid < bsize) {
                const int j = blocks[bi * bsize + tid];
                bs.mem[tid] = j;
                bs.bold[tid] = ((mask_s[j >> 5] >> (j & 31)) & 1u) ? __ldcg(beta_g + j) : 0.0;
            }
            __syncthreads();
            if (!GRAM) {
                for (int k = 0; k < bsize; ++k) {                      // zero the block (:430-442)
                    const double bold = bs.bold[k];
                    if (bold != 0) {                                   // block-uniform
                        const int j = bs.mem[k];
                        add_column(j, bold);
                        if (tid == 0) { __stcg(beta_g + j, 0.0); mask_s[j >> 5] &= ~(1u << (j & 31)); }
                    }
                }
            }
            // sufficient statistics of the zeroed block (:83-160).  Gram form: the block is not
            // zeroed by a pass over q; X_A^T r of the zeroed block follows from q and the block of
            // X^T X alone (q_m + sum_k beta_k G[m][k], the same fma chain as the passes), and the
            // old coefficients are folded into the write-back below -- one Gram column per member
            // whose coefficient actually changes instead of two.
            {
                const double* gb = P.gb + (int64_t)(bs.mem[0] / bsize) * bsize * bsize;
                if (tid < bsize * bsize) bs.A[tid] = gb[tid];
                if (!GRAM && warp < bsize) { const double d = warp_dot(bs.mem[warp]); if (lane == 0) bs.xatr[warp] = d; }
            }
            __syncthreads();
            if (GRAM) {
                if (tid < bsize) {
                    double x = q[bs.mem[tid]];
                    for (int k = 0; k < bsize; ++k) if (bs.bold[k] != 0) x = fma(bs.bold[k], bs.A[tid * bsize + k], x);
                    bs.xatr[tid] = x;
                }
                __syncthreads();
            }
            BPH(3);
            if (warp == 0) {
                bool bad2 = false;
                double denom = 0.0;
                const double* pr = probs;
                if (GRAM && P.npad <= 32) {
                    // the window scored this block on the same (zeroed) statistics: take its model
                    // probabilities and accumulate the normaliser in the order of _weighted_choice
                    const int wf = first / G, gf = first - wf * G;
                    pr = probs_all + wf * nmodel_pad + gf * P.npad;
                    const unsigned gm = (P.npad == 32 ? 0xffffffffu : ((1u << P.npad) - 1u)) << (gf * P.npad);
                    bad2 = (bs.failb[wf] & gm) != 0u;
                    if (lane == 0) for (int ii = 0; ii < P.nmodel; ++ii) denom += pr[ii];
                } else {
                    denom = block_score(P, bs.A, bs.xatr, tau2, sigma2, logp0, log1p0, probs, &bad2);
                }
                if (lane < bsize) bs.zn[lane] = z_s ? z_s[bi * bsize + lane] : rng_normal(key, KIND_BLOCK, bi, sweep, 1 + lane);
                const bool par = GRAM && P.npad <= 32;                  // one quotient per lane, summed in order below
                if (par) {
                    denom = __shfl_sync(0xffffffffu, denom, 0);
                    if (lane < P.nmodel) bs.cq[lane] = pr[lane] / denom;
                }
                __syncwarp();
                if (lane == 0) {
                    const double u = par ? bs.u[first / G][first % G]
                                         : (u_s ? u_s[bi] : rng_uniform(key, KIND_BLOCK, bi, sweep, 0));
                    int mj = 0;
                    double cumsum = 0.0;
                    for (int ii = 0; ii < P.nmodel; ++ii) {            // _weighted_choice :72-77
                        cumsum += par ? bs.cq[ii] : pr[ii] / denom;
                        if (cumsum >= u) { mj = ii; break; }
                    }
                    for (int k = 0; k < bsize; ++k) bs.bnew[k] = 0.0;
                    if (mj > 0 && !bad2) {
                        if (!block_draw(P.model_masks[mj - 1], bsize, bs.A, bs.xatr, tau2, sigma2, bs.zn, bs.bnew)) bad2 = true;
                    }
                    if (bad2) { *P.numeric_err = 1; mj = 0; for (int k = 0; k < bsize; ++k) bs.bnew[k] = 0.0; }
                    bs.mj = mj;
                }
            }
            __syncthreads();
            BPH(4);
            const int mj = bs.mj;
            if (GRAM) {                                                // zeroing and write-back in one update
                const uint32_t msk = mj > 0 ? P.model_masks[mj - 1] : 0u;
                for (int k = 0; k < bsize; ++k) {
                    const double bold = bs.bold[k];
                    const bool sel = (msk >> k) & 1u;
                    if (bold == 0 && !sel) continue;                   // block-uniform
                    const int j = bs.mem[k];
                    const double bnew = sel ? bs.bnew[k] : 0.0;
                    if (bnew != bold) add_column(j, bold - bnew);
                    if (tid == 0) {
                        __stcg(beta_g + j, bnew);
                        if (bnew != 0) mask_s[j >> 5] |= (1u << (j & 31));
                        else mask_s[j >> 5] &= ~(1u << (j & 31));
                    }
                }
            } else if (mj > 0) {                                       // write back (:850-884)
                const uint32_t msk = P.model_masks[mj - 1];
                for (int k = 0; k < bsize; ++k) {
                    if (!(msk & (1u << k))) continue;
                    const int j = bs.mem[k];
                    const double bnew = bs.bnew[k];
                    add_column(j, -bnew);
                    if (tid == 0) {
                        __stcg(beta_g + j, bnew);
                        if (bnew != 0) mask_s[j >> 5] |= (1u << (j & 31));
                        else mask_s[j >> 5] &= ~(1u << (j & 31));
                    }
                }
            }
            if (mj > 0) {
                if (PROBIT) {
                    for (int k = tid; k < n; k += GB_BLOCK) {
                        const double m = mu_s[k];
                        const double yv = truncnorm(key, KIND_TRUNCNORM, (uint32_t)k, event, m, P.cfg.sigma2, 0.0, P.zlab[k]);
                        y_s[k] = yv;
                        r_s[k] = yv - m;
                    }
                    event++;
                    dirty = true;
                }
            }
            __syncthreads();
            BPH(5);
            pos0 = bi + 1;
        }

        double ssr = 0.0;
        if (GRAM) ssr = sh.rr;
        else {
            for (int k = tid; k < n; k += GB_BLOCK) ssr = fma(r_s[k], r_s[k], ssr);
            ssr = block_sum(ssr, scratch);
        }
        sweep_epilogue(P, &sh, mask_s, beta_g, scratch, key, c, s, ssr, y_s, dirty);
        dirty = false;
        __syncthreads();
        BPH(6);
        BPH_PRINT;
    }

    if (GRAM) for (int k = tid; k < p2; k += GB_BLOCK) P.q[(int64_t)c * p2 + k] = q[k];
    else for (int k = tid; k < n2; k += GB_BLOCK) {
        P.r[(int64_t)c * n2 + k] = r_s[k];
        if (PROBIT) { P.mu[(int64_t)c * n2 + k] = mu_s[k]; P.ylat[(int64_t)c * n2 + k] = y_s[k]; }
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2;
        P.hstate[GB_HS * c + 2] = sh.p0; P.hstate[GB_HS * c + 3] = sh.rr;
        P.hstate[GB_HS * c + 4] += (double)n_changes; P.hstate[GB_HS * c + 5] += (double)n_windows;
        if (PROBIT) { P.events[2 * c] = event; P.events[2 * c + 1] = dirty ? 1u : 0u; }
    }
}
