// nprior.cu -- multi-chain neuronized-prior Gibbs sampler (beta_j = w_j * relu(alpha_j - alpha0)).
//
// Replaces `apply_pool(_nprior_sample, ...)` (/root/reference/pyblip/nprior/nprior.py:126-153)
// and the Cython kernel /root/reference/pyblip/nprior/_nprior.pyx:56-678.
//
// One CTA per chain, Gram form: the reference's three O(n) passes per coordinate
// (dnrm2, ddot, dnrm2 on r + alpha0*w_j*X_j; _nprior.pyx:617-652) are O(1) given
// q = X^T r and rr = ||r||^2:
//     X_j.(r + s X_j) = q_j + s G_jj,        ||r + s X_j||^2 = rr + 2 s q_j + s^2 G_jj.
// A visit that leaves the coordinate inactive (alpha_j <= alpha0 before and after) does not
// touch r, so the block evaluates a window of NP_BLOCK upcoming visits in parallel, one per
// thread and active coordinates included (exact speculation, as in gibbs_kernels.cuh), commits
// the leading visits that stay inactive and applies the first one that touches r as evaluated,
// with one Gram-column axpy by the whole block.  alpha, w, q and ||X_j||^2 live in dynamic shared
// memory when 4 p doubles fit; the state-independent variates of a visit are drawn once per sweep.
// Every sweep starts from fresh w (joint draw or the reference's fresh 0.1*N(0,1) row,
// :100,132-140), so q and rr are rebuilt from X^T y at every sweep and never drift.  The joint
// draw factors the active-set precision in shared memory (<= 64 active coordinates) or with a
// blocked Cholesky in a per-chain global workspace that grows on demand (launch replayed).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "gibbs_kernels.cuh"

enum : uint32_t { NP_INIT_ALPHA = 16, NP_W_FRESH = 17, NP_SIGMA_INIT = 18, NP_U = 19, NP_TN = 20,
                  NP_WZ = 21, NP_JOINT_ALL = 22, NP_JOINT_ACT = 23, NP_DELTA = 24 };

struct blipgpu_nprior {
    blipgpu_ctx* ctx;
    int64_t rows, p, chains, n_keep;
    double* dense[3];     // alphas, ws, betas: [rows][p]
    double* scal;         // [3][rows]: sigma2s, alpha0s, p0s
    double sweep_ms;
};

namespace {

constexpr int NP_BLOCK = 256;
constexpr int NP_QT = 8;         // coordinates of q a thread keeps in registers while q is rebuilt
constexpr int NP_AMAX_MIN = 512;   // the joint-W Cholesky workspace holds at least this many active coordinates

// variates of one coordinate visit that do not depend on the state: spike uniform, normal of the w
// draw, first normal of the truncated-normal rejection loop (sub-stream NP_TN, attempt 0)
struct __align__(32) NPDraws { double u, z, tn0, pad; };

struct NPParams {
    const double* G; int64_t ldg; const double* Xl2; const double* q0; double yy;
    int n, p, p2;
    blipgpu_nprior_config cfg;
    double min_alpha0, sigma_a;
    // chain state
    double *alpha, *w, *beta;      // [chains][p]
    double* q;                     // [chains][p2]
    NPDraws* uz;                   // [chains][p]: this sweep's state-independent variates by visiting position
    double* hs;                    // [chains][4]: sigma2, alpha0, p0, rr
    int chain_base;                // first chain of this launch (waves)
    int state_smem;                // alpha, w, q and ||X_j||^2 live in dynamic shared memory during the launch
    int amax;                      // capacity of the active-set workspace (<= p, sized from free memory)
    double* chol;                  // [chains][amax*amax]
    double* vec;                   // [chains][3*amax]
    int* act;                      // [chains][amax]
    int* err;                      // 1: active set too large, 2: Cholesky failed
    // streams
    uint32_t k0, k1, chain_offset;
    const void* perm; int perm16, perm_S, perm_s0;
    const double *u, *wz, *alpha_init, *w_fresh, *sigma_init, *joint_all, *joint_act, *delta_z,
                 *invgamma, *p0_draw;            // injected (device, full length) or NULL
    int n_total;
    int sweep0, S, burn, n_keep;
    double* out[3]; double* scal; int64_t rows;
};

__device__ __forceinline__ double rationalapprox_d(double t)
{
    double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328, d0 = 1.432788, d1 = 0.189269, d2 = 0.001308;
    double frac = (c2 * t + c1) * t + c0;
    frac = frac / (((d2 * t + d1) * t + d0) * t + 1.0);
    return t - frac;
}
__host__ __device__ inline double norm_quantile_apprx_hd(double p)
{
    double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328, d0 = 1.432788, d1 = 0.189269, d2 = 0.001308;
    double t = p < 0.5 ? sqrt(-2.0 * log(p)) : sqrt(-2.0 * log(1 - p));
    double frac = (c2 * t + c1) * t + c0;
    frac = frac / (((d2 * t + d1) * t + d0) * t + 1.0);
    return p < 0.5 ? -1 * (t - frac) : (t - frac);
}
__device__ __forceinline__ double normcdf_d(double z) { return 0.5 * erfc(-z * sqrt(0.5)); }

// One coordinate's full update in Gram form (_update_coordinate_relu + _update_wj).
// Inputs: state of the coordinate BEFORE its visit; qj, rr of the current residual.
struct NPVisit { double alpha_new, w_new, beta_new, qj_back, rr_back; bool stays_inactive; };

__device__ inline NPVisit np_visit(const NPParams& P, const RngKey& key, int pos, int sweep, double u, double z,
                                   double tn0, double log_p0,
                                   double alpha_old, double w_old, double xl2, double qj, double rr,
                                   double sigma2, double alpha0, double p0)
{
    NPVisit v;
    const double b_old = w_old * fmax(alpha_old - alpha0, 0.0);
    // add back: r' = r + b_old X_j
    const double qb = qj + b_old * xl2;
    const double rrb = rr + 2.0 * b_old * qj + b_old * b_old * xl2;
    v.qj_back = qb; v.rr_back = rrb;
    // _update_coordinate_relu (:593-678)
    const double wj = w_old;
    double r2 = rrb;
    const double r_scale = alpha0 * wj;
    const double t_denom = wj * wj * xl2 + sigma2;
    const double t_sigma2 = sigma2 / t_denom;
    double t_alphaj = qb + r_scale * xl2;                       // X_j . (r' + r_scale X_j)
    t_alphaj = wj * t_alphaj / t_denom;
    const double rplus2 = rrb + 2.0 * r_scale * qb + r_scale * r_scale * xl2;
    double log_kappa_num = log_p0;
    log_kappa_num -= r2 / (2 * sigma2);
    double log_kappa_denom = log(1 - normcdf_d((alpha0 - t_alphaj) / sqrt(t_sigma2)));
    log_kappa_denom += log(t_sigma2) / 2 + (t_alphaj * t_alphaj / (2 * t_sigma2) - rplus2 / (2 * sigma2));
    const double ratio = exp(log_kappa_num - log_kappa_denom);
    const double kappa = ratio / (1.0 + ratio);
    if (u < kappa) v.alpha_new = truncnorm(key, NP_TN, pos, sweep, 0.0, 1.0, alpha0, 1, &tn0);
    else v.alpha_new = truncnorm(key, NP_TN, pos, sweep, t_alphaj, t_sigma2, alpha0, 0, &tn0);
    // _update_wj (:543-586)
    const double Tj = fmax(v.alpha_new - alpha0, 0.0);
    const double sigma02 = P.cfg.sigma_prior_type == 0 ? 1.0 : sigma2;
    const double vj = xl2 * Tj * Tj + (sigma02 / P.cfg.tauw2);
    double mj = 0.0;
    if (Tj > 0) mj = qb * Tj / vj;
    v.w_new = sqrt(sigma2 / vj) * z + mj;
    v.beta_new = v.w_new * Tj;
    v.stays_inactive = (b_old == 0.0) && (v.beta_new == 0.0);
    return v;
}

// In-place lower Cholesky of the A x A matrix L (row-major, lower triangle valid) by one CTA,
// right-looking with 32-wide panels: the panel is factored / solved against a 32 x 32 diagonal
// block held in shared memory, and the trailing matrix is updated tile by tile (64 x 64 tiles,
// 4 x 4 outputs per thread, both panel strips in shared memory), so that the trailing matrix is
// read and written once per PANEL instead of once per column.  Used for large active sets, where
// the column-by-column version streams A^3/6 elements through L2.
constexpr int NPC_NB = 32, NPC_T = 64;
struct NpCholSmem { double D[NPC_NB][NPC_NB + 1]; double Pr[NPC_T][NPC_NB + 1]; double Pc[NPC_T][NPC_NB + 1]; };

// Active sets of at most NPS_A coordinates (the steady state of a sparse posterior) are factored
// and solved entirely in shared memory; the storage is shared with the blocked factorisation.
constexpr int NPS_A = 64;
struct NpSmallSmem { double L[NPS_A][NPS_A + 1]; double tmu[NPS_A]; double wt[NPS_A]; };
union NpSmem { NpCholSmem blocked; NpSmallSmem small; };

__device__ inline void np_cholesky_blocked(double* L, int A, int* err, NpCholSmem* sm)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int kb = 0; kb < A; kb += NPC_NB) {
        const int nb = min(NPC_NB, A - kb);
        // (1) diagonal block: load, factor (warp 0, lane = row), store
        for (int e = tid; e < nb * nb; e += NP_BLOCK) {
            const int r = e / nb, c2 = e % nb;
            sm->D[r][c2] = c2 <= r ? L[(int64_t)(kb + r) * A + kb + c2] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            for (int k = 0; k < nb; ++k) {
                const double dk = sm->D[k][k];
                if (lane == 0 && !(dk > 0)) *err = 2;
                const double d = sqrt(dk);
                __syncwarp();
                if (lane == k) sm->D[k][k] = d;
                if (lane > k && lane < nb) sm->D[lane][k] /= d;
                __syncwarp();
                if (lane > k && lane < nb) {
                    const double lrk = sm->D[lane][k];
                    for (int c2 = k + 1; c2 <= lane; ++c2) sm->D[lane][c2] -= lrk * sm->D[c2][k];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += NP_BLOCK) {
            const int r = e / nb, c2 = e % nb;
            if (c2 <= r) L[(int64_t)(kb + r) * A + kb + c2] = sm->D[r][c2];
        }
        // (2) panel below the block: row r of the panel solves x D^T = a (one thread per row)
        const int r0 = kb + nb;
        for (int r = r0 + tid; r < A; r += NP_BLOCK) {
            double* row = L + (int64_t)r * A + kb;
            double x[NPC_NB];
#pragma unroll
            for (int j = 0; j < NPC_NB; ++j) x[j] = j < nb ? row[j] : 0.0;
#pragma unroll
            for (int j = 0; j < NPC_NB; ++j) {
                if (j < nb) {
                    double sacc = x[j];
#pragma unroll
                    for (int i = 0; i < NPC_NB; ++i) if (i < j) sacc -= x[i] * sm->D[j][i];
                    x[j] = sacc / sm->D[j][j];
                }
            }
#pragma unroll
            for (int j = 0; j < NPC_NB; ++j) if (j < nb) row[j] = x[j];
        }
        __syncthreads();
        // (3) trailing update L[r][c] -= sum_j P[r][j] P[c][j], r >= c >= r0, 64 x 64 tiles.
        // A thread owns the 4 x 4 outputs (tr + 16 a, tc + 16 b): the 16 lanes of a half-warp read 16
        // different panel rows (row stride 33 doubles: conflict-free) and write 128-byte row segments.
        // The outputs are loaded into the accumulators before the panel strip is staged, and the next
        // strip is fetched into registers while this one is multiplied, so global latency is covered.
        const int tr = tid >> 4, tc = tid & 15;
        constexpr int NPF = NPC_T * NPC_NB / NP_BLOCK;                 // strip elements a thread stages
        for (int rb = r0; rb < A; rb += NPC_T) {
            __syncthreads();
            for (int e = tid; e < NPC_T * NPC_NB; e += NP_BLOCK) {    // panel strip of the row block
                const int r = e / NPC_NB, j = e % NPC_NB;
                sm->Pr[r][j] = (rb + r < A && j < nb) ? L[(int64_t)(rb + r) * A + kb + j] : 0.0;
            }
            double nxt[NPF];
#pragma unroll
            for (int i = 0; i < NPF; ++i) {
                const int e = tid + i * NP_BLOCK, r = e / NPC_NB, j = e % NPC_NB;
                nxt[i] = (r0 + r < A && j < nb) ? L[(int64_t)(r0 + r) * A + kb + j] : 0.0;
            }
            for (int cb = r0; cb <= rb; cb += NPC_T) {
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = rb + tr + 16 * a;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int c2 = cb + tc + 16 * b;
                        acc[a][b] = (r < A && c2 <= r) ? L[(int64_t)r * A + c2] : 0.0;
                    }
                }
                __syncthreads();                                      // previous tile done with Pc
#pragma unroll
                for (int i = 0; i < NPF; ++i) { const int e = tid + i * NP_BLOCK; sm->Pc[e / NPC_NB][e % NPC_NB] = nxt[i]; }
                __syncthreads();
                if (cb + NPC_T <= rb) {
#pragma unroll
                    for (int i = 0; i < NPF; ++i) {
                        const int e = tid + i * NP_BLOCK, r = e / NPC_NB, j = e % NPC_NB;
                        nxt[i] = (cb + NPC_T + r < A && j < nb) ? L[(int64_t)(cb + NPC_T + r) * A + kb + j] : 0.0;
                    }
                }
#pragma unroll 8
                for (int j = 0; j < NPC_NB; ++j) {
                    double pr[4], pc[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { pr[a] = sm->Pr[tr + 16 * a][j]; pc[a] = sm->Pc[tc + 16 * a][j]; }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(-pr[a], pc[b], acc[a][b]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = rb + tr + 16 * a;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int c2 = cb + tc + 16 * b;
                        if (r < A && c2 <= r) L[(int64_t)r * A + c2] = acc[a][b];
                    }
                }
            }
        }
        __syncthreads();
    }
}

#ifdef NP_PHASE_CLOCKS
#define NPH_DECL long long ph_t[12] = {0,0,0,0,0,0,0,0,0,0,0,0}; long long ph_last = clock64(); int ph_iter = 0;
#define NPH(i) do { if (threadIdx.x == 0) { long long _n = clock64(); ph_t[i] += _n - ph_last; ph_last = _n; } } while (0)
#define NPH_ITER (++ph_iter)
#define NPH_PRINT do { if (threadIdx.x == 0 && ((blockIdx.x == 0 && (sweep < 6 || sweep % 50 == 0)) || A > 300 || ph_iter > 400)) \
    printf("chain %d sweep %d A %d iters %d | draws %lld ptp %lld chol %lld solve %lld qbuild %lld | eval %lld find %lld commit %lld qupd %lld rest %lld | tail %lld\n", (int)blockIdx.x, sweep, A, ph_iter, \
           ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[6], ph_t[7], ph_t[8], ph_t[9], ph_t[10]); printf("   fwd(incl. 2 div) %lld\n", ph_t[11]); } while (0)
#else
#define NPH_DECL
#define NPH(i) do {} while (0)
#define NPH_ITER do {} while (0)
#define NPH_PRINT do {} while (0)
#endif

__global__ void __launch_bounds__(NP_BLOCK, 2) nprior_kernel(NPParams P)
{
    __shared__ double scratch[40];
    __shared__ double s_sigma2, s_alpha0, s_p0, s_rr, s_delta, s_w;
    __shared__ int s_j, s_A;
    __shared__ unsigned s_flags[NP_BLOCK / 32];
    __shared__ NpSmem s_mem;
    NpCholSmem& s_chol = s_mem.blocked;
    __shared__ int s_scan[NP_BLOCK / 32 + 2];

    // chain c of this rank; the joint-draw workspace is indexed by the slot within the launch, because
    // the chains run in waves when the workspace of all of them does not fit in memory
    const int slot = blockIdx.x, c = P.chain_base + slot, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = P.p, p2 = P.p2;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    // chain state: in dynamic shared memory for the whole launch when it fits (4 p doubles), else in
    // the chain's global arrays; the code below only sees the pointers
    extern __shared__ double np_dyn[];
    double* const g_alpha = P.alpha + (int64_t)c * p;
    double* const g_w = P.w + (int64_t)c * p;
    double* const g_q = P.q + (int64_t)c * p2;
    double* alpha = g_alpha; double* w = g_w; double* q = g_q;
    const double* Xl2 = P.Xl2;
    if (P.state_smem) {
        alpha = np_dyn; w = np_dyn + p2; q = np_dyn + 2 * p2;
        double* xs = np_dyn + 3 * p2;
        for (int k = tid; k < p; k += NP_BLOCK) { alpha[k] = g_alpha[k]; w[k] = g_w[k]; q[k] = g_q[k]; xs[k] = P.Xl2[k]; }
        for (int k = p + tid; k < p2; k += NP_BLOCK) q[k] = 0.0;
        Xl2 = xs;
    }
    double* beta = P.beta + (int64_t)c * p;
    NPDraws* uz = P.uz + (int64_t)c * p;
    const int NP_AMAX = P.amax;
    double* L = P.chol + (int64_t)slot * NP_AMAX * NP_AMAX;
    double* vec = P.vec + (int64_t)slot * 3 * NP_AMAX;
    int* act = P.act + (int64_t)slot * NP_AMAX;
    const blipgpu_nprior_config& cf = P.cfg;

    if (tid == 0) { s_sigma2 = P.hs[4 * c]; s_alpha0 = P.hs[4 * c + 1]; s_p0 = P.hs[4 * c + 2]; s_rr = P.hs[4 * c + 3]; }
    __syncthreads();

    for (int s = 0; s < P.S; ++s) {
        const int sweep = P.sweep0 + s;
        NPH_DECL
        const int64_t soff = ((int64_t)c * P.n_total + sweep) * p;      // injected arrays: full length
        const PermView perm_s = { P.perm, ((int64_t)c * P.perm_S + P.perm_s0 + s) * p, P.perm16 };
        const double sigma2 = s_sigma2, alpha0 = s_alpha0, p0 = s_p0;

        // the state-independent variates of the coordinate visits (spike uniform, normal of the w
        // draw) are keyed by position: drawn once per sweep here, not at each re-evaluation
        for (int pos = tid; pos < p; pos += NP_BLOCK) {
            NPDraws d;
            d.u = P.u ? P.u[soff + pos] : rng_uniform(key, NP_U, pos, sweep, 0);
            d.z = P.wz ? P.wz[soff + pos] : rng_normal(key, NP_WZ, pos, sweep, 0);
            d.tn0 = rng_normal(key, NP_TN, pos, sweep, 0);
            d.pad = 0.0;
            uz[pos] = d;
        }
        const double log_p0 = log(p0);
        // ---- step 1: w of this sweep (:132-140 joint draw, or the fresh 0.1*N(0,1) row :100)
        int nact = 0;
        for (int j0 = 0; j0 < p; j0 += NP_BLOCK) {                     // ordered compaction of the active set
            const int j = j0 + tid;
            bool a = false;
            if (j < p) {
                a = fmax(alpha[j] - alpha0, 0.0) > 0;
                if (cf.joint_sample_W == 1) {
                    const double z = P.joint_all ? P.joint_all[soff + j] : rng_normal(key, NP_JOINT_ALL, j, sweep, 0);
                    w[j] = z / sqrt(sigma2 / cf.tauw2);
                } else {
                    const double z = P.w_fresh ? P.w_fresh[soff + j] : rng_normal(key, NP_W_FRESH, j, sweep, 0);
                    w[j] = 0.1 * z;
                }
            }
            int total;
            const int pre = block_excl_scan(a ? 1 : 0, s_scan, &total);
            if (a && nact + pre < NP_AMAX) act[nact + pre] = j;
            nact += total;
        }
        if (tid == 0) s_A = nact;
        __syncthreads();
        const int A = nact;
        // another chain outgrew the workspace: this launch is replayed.  Report this chain's own
        // active-set size first, so that one replay sizes the workspace for all chains.
        if (*reinterpret_cast<volatile int*>(P.err) == 1) { if (tid == 0) atomicMax(P.err + 1, A); return; }
        NPH(0);
        if (cf.joint_sample_W == 1 && A > 0) {
            if (A > NP_AMAX) { if (tid == 0) { atomicMax(P.err + 1, A); *P.err = 1; } return; }
            if (A <= NPS_A) {
                // small active set: precision, factor and both solves in shared memory.  Same
                // operation order per element as the global-memory path below.
                NpSmallSmem& S = s_mem.small;
                for (int e = tid; e < A * A; e += NP_BLOCK) {
                    const int r = e / A, cc = e % A;
                    if (cc <= r) {
                        const int jr = act[r], jc = act[cc];
                        double v = fmax(alpha[jr] - alpha0, 0.0) * fmax(alpha[jc] - alpha0, 0.0) * P.G[(int64_t)jr * P.ldg + jc];
                        if (r == cc) v += 1 / cf.tauw2;
                        S.L[r][cc] = v;
                    }
                }
                for (int a = tid; a < A; a += NP_BLOCK) {
                    const int j = act[a];
                    S.tmu[a] = fmax(alpha[j] - alpha0, 0.0) * P.q0[j];                  // Phi y
                    S.wt[a] = P.joint_act ? P.joint_act[soff + a] : rng_normal(key, NP_JOINT_ACT, a, sweep, 0);
                }
                __syncthreads();
                NPH(1);
                // right-looking factorisation with ONE barrier per column: column k stays unscaled
                // (L[r][k] sqrt(d_k)) while the trailing update uses L[r][k] L[c][k] / d_k, and all
                // columns are scaled by 1 / sqrt(d_k) in one pass at the end
                for (int k = 0; k < A; ++k) {
                    const double d = S.L[k][k];
                    if (tid == 0 && !(d > 0)) *P.err = 2;
                    const double dinv = 1.0 / d;
                    for (int r = k + 1 + warp; r < A; r += NP_BLOCK / 32) {
                        const double lrk = S.L[r][k] * dinv;
                        for (int cc = k + 1 + lane; cc <= r; cc += 32) S.L[r][cc] = fma(-lrk, S.L[cc][k], S.L[r][cc]);
                    }
                    __syncthreads();
                }
                for (int e = tid; e < A * A; e += NP_BLOCK) {
                    const int r = e / A, cc = e % A;
                    if (cc < r) S.L[r][cc] *= rsqrt(S.L[cc][cc]);
                }
                __syncthreads();
                if (tid < A) S.L[tid][tid] = sqrt(S.L[tid][tid]);
                __syncthreads();
                NPH(2);
                if (warp == 0) {
                    // triangular solves in column (axpy) form, lane = row (rows lane and lane + 32):
                    // forward L v = tmu, then backward L^T m = v and L^T s = z together
                    const unsigned full = 0xffffffffu;
                    const int r0 = lane, r1 = lane + 32;
                    const double id0 = r0 < A ? 1.0 / S.L[r0][r0] : 0.0, id1 = r1 < A ? 1.0 / S.L[r1][r1] : 0.0;
                    double t0 = r0 < A ? S.tmu[r0] : 0.0, t1 = r1 < A ? S.tmu[r1] : 0.0;
                    for (int k = 0; k < A; ++k) {
                        const double xk = __shfl_sync(full, k < 32 ? t0 * id0 : t1 * id1, k & 31);
                        if (r0 == k) t0 = xk;
                        if (r1 == k) t1 = xk;
                        if (r0 > k && r0 < A) t0 = fma(-S.L[r0][k], xk, t0);
                        if (r1 > k && r1 < A) t1 = fma(-S.L[r1][k], xk, t1);
                    }
                    NPH(11);
                    double z0 = r0 < A ? S.wt[r0] : 0.0, z1 = r1 < A ? S.wt[r1] : 0.0;
                    for (int k = A - 1; k >= 0; --k) {
                        const double mk = __shfl_sync(full, k < 32 ? t0 * id0 : t1 * id1, k & 31);
                        const double sk = __shfl_sync(full, k < 32 ? z0 * id0 : z1 * id1, k & 31);
                        if (r0 == k) { t0 = mk; z0 = sk; }
                        if (r1 == k) { t1 = mk; z1 = sk; }
                        if (r0 < k) { const double l = S.L[k][r0]; t0 = fma(-l, mk, t0); z0 = fma(-l, sk, z0); }
                        if (r1 < k) { const double l = S.L[k][r1]; t1 = fma(-l, mk, t1); z1 = fma(-l, sk, z1); }
                    }
                    if (r0 < A) { S.tmu[r0] = t0; S.wt[r0] = z0; }
                    if (r1 < A) { S.tmu[r1] = t1; S.wt[r1] = z1; }
                }
                __syncthreads();
                for (int a = tid; a < A; a += NP_BLOCK) w[act[a]] = S.wt[a] + S.tmu[a];
                __syncthreads();
            } else {
                // large active set (early sweeps): precision PTP = T G_AA T + I / tauw2 (:307-337),
                // lower triangle, in the chain's global workspace; blocked factorisation
                for (int e = tid; e < A * A; e += NP_BLOCK) {
                    const int r = e / A, cc = e % A;
                    if (cc <= r) {
                        const int jr = act[r], jc = act[cc];
                        double v = fmax(alpha[jr] - alpha0, 0.0) * fmax(alpha[jc] - alpha0, 0.0) * P.G[(int64_t)jr * P.ldg + jc];
                        if (r == cc) v += 1 / cf.tauw2;
                        L[(int64_t)r * A + cc] = v;
                    }
                }
                __syncthreads();
                NPH(1);
                np_cholesky_blocked(L, A, P.err, &s_chol);
                NPH(2);
                // mean and draw by triangular solves (:371-432)
                double* tmu = vec; double* wt = vec + NP_AMAX;
                for (int a = tid; a < A; a += NP_BLOCK) {
                    const int j = act[a];
                    tmu[a] = fmax(alpha[j] - alpha0, 0.0) * P.q0[j];                    // Phi y
                    wt[a] = P.joint_act ? P.joint_act[soff + a] : rng_normal(key, NP_JOINT_ACT, a, sweep, 0);
                }
                __syncthreads();
                // the whole block works on one row of L at a time (contiguous loads).
                // forward  L v = tmu  by rows, with the partial sums reduced over the block;
                // backward L^T m = v, L^T s = z  in axpy form: once x_r is known, row r of L updates
                // the right-hand sides of all m < r.
                for (int r = 0; r < A; ++r) {
                    double acc = 0.0;
                    for (int m = tid; m < r; m += NP_BLOCK) acc = fma(L[(int64_t)r * A + m], tmu[m], acc);
                    acc = block_sum(acc, scratch);
                    if (tid == 0) tmu[r] = (tmu[r] - acc) / L[(int64_t)r * A + r];
                    __syncthreads();
                }
                for (int r = A - 1; r >= 0; --r) {
                    const double lrr = L[(int64_t)r * A + r];
                    const double xm = tmu[r] / lrr, xs = wt[r] / lrr;
                    __syncthreads();
                    if (tid == 0) { tmu[r] = xm; wt[r] = xs; }
                    for (int m = tid; m < r; m += NP_BLOCK) {
                        const double l = L[(int64_t)r * A + m];
                        tmu[m] = fma(-l, xm, tmu[m]); wt[m] = fma(-l, xs, wt[m]);
                    }
                    __syncthreads();
                }
                for (int a = tid; a < A; a += NP_BLOCK) w[act[a]] = wt[a] + tmu[a];
                __syncthreads();
            }
        }

        NPH(3);
        // ---- beta, q = X^T(y - X beta), rr  (:143-160)
        for (int j = tid; j < p; j += NP_BLOCK) beta[j] = w[j] * fmax(alpha[j] - alpha0, 0.0);
        __syncthreads();
        if (A <= NP_AMAX) {
            // each thread carries NP_QT of its coordinates in registers through the active columns
            // (ascending, the order of the column-by-column form), so the loads are independent
            for (int kt = 0; kt < p2; kt += NP_QT * NP_BLOCK) {
                double acc[NP_QT];
#pragma unroll
                for (int i = 0; i < NP_QT; ++i) { const int k = kt + tid + i * NP_BLOCK; acc[i] = k < p ? P.q0[k] : 0.0; }
#pragma unroll 2
                for (int a = 0; a < A; ++a) {
                    const int j = act[a];
                    const double bj = beta[j];
                    const double* g = P.G + (int64_t)j * P.ldg + kt + tid;
#pragma unroll
                    for (int i = 0; i < NP_QT; ++i) if (kt + tid + i * NP_BLOCK < p) acc[i] = fma(-bj, g[i * NP_BLOCK], acc[i]);
                }
#pragma unroll
                for (int i = 0; i < NP_QT; ++i) { const int k = kt + tid + i * NP_BLOCK; if (k < p2) q[k] = acc[i]; }
            }
        } else {                                    // joint draw off and a large active set: scan all
            for (int k = tid; k < p2; k += NP_BLOCK) q[k] = k < p ? P.q0[k] : 0.0;
            __syncthreads();
            for (int j = 0; j < p; ++j) {
                const double bj = beta[j];
                if (bj != 0) for (int k = tid; k < p; k += NP_BLOCK) q[k] = fma(-bj, P.G[(int64_t)j * P.ldg + k], q[k]);
            }
        }
        __syncthreads();
        {
            double acc = 0.0;
            for (int j = tid; j < p; j += NP_BLOCK) { const double bj = beta[j]; if (bj != 0) acc = fma(bj, P.q0[j] + q[j], acc); }
            acc = block_sum(acc, scratch);
            if (tid == 0) s_rr = P.yy - acc;
        }
        __syncthreads();

        // ---- coordinate loop (:163-198): a window of NP_BLOCK upcoming positions, one per thread, is
        // evaluated against the current state; the leading visits that stay inactive are committed
        // (they redraw alpha_j and w_j but leave r alone) and the first one that touches r is applied
        NPH(4);
        int pos0 = 0;
        while (pos0 < p) {
            NPH_ITER;
            const int T = min(NP_BLOCK, p - pos0);
            NPVisit v; v.stays_inactive = false;
            int j = 0; double a_old = 0, w_old = 0, xl2 = 0, qj = 0;
            bool change = false;
            if (tid < T) {
                j = perm_s[pos0 + tid];
                a_old = alpha[j]; w_old = w[j]; xl2 = Xl2[j]; qj = q[j];
                // every visit is evaluated against the current state, active coordinates included: the
                // first one that touches r is then applied as evaluated (active before => never inactive)
                const NPDraws d = uz[pos0 + tid];
                v = np_visit(P, key, pos0 + tid, sweep, d.u, d.z, d.tn0, log_p0, a_old, w_old, xl2, qj, s_rr, sigma2, alpha0, p0);
                change = !v.stays_inactive;
            }
            NPH(5);
            const unsigned m = __ballot_sync(0xffffffffu, change);
            if (lane == 0) s_flags[warp] = m;
            __syncthreads();
            int first = -1;
#pragma unroll
            for (int wv = NP_BLOCK / 32 - 1; wv >= 0; --wv) if (s_flags[wv]) first = 32 * wv + __ffs(s_flags[wv]) - 1;
            NPH(6);
            if (tid < T && (first < 0 || tid < first)) { alpha[j] = v.alpha_new; w[j] = v.w_new; }
            if (first >= 0 && tid == first) {
                const double b_old = w_old * fmax(a_old - alpha0, 0.0);
                alpha[j] = v.alpha_new; w[j] = v.w_new; beta[j] = v.beta_new;
                s_rr = v.rr_back - 2.0 * v.beta_new * v.qj_back + v.beta_new * v.beta_new * xl2;
                s_delta = v.beta_new - b_old; s_j = j;
            }
            __syncthreads();
            NPH(7);
            if (first < 0) { pos0 += T; continue; }
            const double delta = s_delta; const int jstar = s_j;
            if (delta != 0) {
                // the column is loaded into registers first: a store to q between two loads of G would
                // serialise the round trips (q and G may alias as far as the compiler knows)
                const double* gcol = P.G + (int64_t)jstar * P.ldg;
                for (int kt = 0; kt < p; kt += NP_QT * NP_BLOCK) {
                    double g[NP_QT];
#pragma unroll
                    for (int i = 0; i < NP_QT; ++i) { const int k = kt + tid + i * NP_BLOCK; g[i] = k < p ? __ldg(gcol + k) : 0.0; }
#pragma unroll
                    for (int i = 0; i < NP_QT; ++i) { const int k = kt + tid + i * NP_BLOCK; if (k < p) q[k] = fma(-delta, g[i], q[k]); }
                }
            }
            __syncthreads();
            NPH(8);
            pos0 += first + 1;
        }

        NPH(9);
        // ---- sigma2 (:510-537)
        {
            double w2 = 0.0;
            if (cf.sigma_prior_type != 0) {
                for (int j = tid; j < p; j += NP_BLOCK) w2 = fma(w[j], w[j], w2);
                w2 = block_sum(w2, scratch);
                w2 = sqrt(w2); w2 = w2 * w2;
            }
            if (tid == 0) {
                double r2 = sqrt(s_rr); r2 = r2 * r2;
                double b = r2 / 2 + cf.sigma_b0;
                if (cf.sigma_prior_type != 0) b = b + w2 / (2.0 * cf.tauw2);
                const double ig = P.invgamma ? P.invgamma[(int64_t)c * P.n_total + sweep]
                                             : 1.0 / rng_gamma(key, SLOT_INVGAMMA, sweep, P.sigma_a);
                s_sigma2 = b * ig;
            }
        }
        // ---- sparsity (:218-236)
        if (cf.update_p0 != 0) {
            if (cf.group_alpha_update == 0) {
                int k = 0;
                for (int j = tid; j < p; j += NP_BLOCK) k += (alpha[j] > alpha0) ? 1 : 0;
                k = (int)(block_sum((double)k, scratch) + 0.5);
                if (tid == 0) {
                    const double v = P.p0_draw ? P.p0_draw[(int64_t)c * P.n_total + sweep]
                                               : rng_beta(key, 0, sweep, 1.0 + p - k, 1.0 + k);
                    s_p0 = fmax(cf.min_p0, v);
                    s_alpha0 = norm_quantile_apprx_hd(s_p0);
                }
                __syncthreads();
            } else {
                double sum = 0.0;
                for (int j = tid; j < p; j += NP_BLOCK) sum += alpha[j];
                sum = block_sum(sum, scratch);
                if (tid == 0) {
                    const double pf = p;
                    const double mean = -1 * (sum + alpha0) / (pf + 1);
                    const double z = P.delta_z ? P.delta_z[(int64_t)c * P.n_total + sweep] : rng_normal(key, NP_DELTA, 0, sweep, 0);
                    double delta = 1 / sqrt(pf + 1) * z + mean;
                    delta = fmax(P.min_alpha0 - alpha0, delta);
                    s_w = delta;
                    s_alpha0 = alpha0 + delta;
                    s_p0 = normcdf_d(s_alpha0);
                }
                __syncthreads();
                const double delta = s_w;
                for (int j = tid; j < p; j += NP_BLOCK) alpha[j] += delta;
            }
        }
        __syncthreads();
        // ---- record
        if (sweep >= P.burn) {
            const int64_t row = (int64_t)c * P.n_keep + (sweep - P.burn);
            for (int j = tid; j < p; j += NP_BLOCK) {
                P.out[0][row * p + j] = alpha[j];
                P.out[1][row * p + j] = w[j];
                P.out[2][row * p + j] = beta[j];
            }
            if (tid == 0) {
                P.scal[row] = s_sigma2; P.scal[P.rows + row] = s_alpha0; P.scal[2 * P.rows + row] = s_p0;
            }
        }
        __syncthreads();
        NPH(10);
        NPH_PRINT;
    }
    if (tid == 0) { P.hs[4 * c] = s_sigma2; P.hs[4 * c + 1] = s_alpha0; P.hs[4 * c + 2] = s_p0; P.hs[4 * c + 3] = s_rr; }
    if (P.state_smem)
        for (int k = tid; k < p; k += NP_BLOCK) { g_alpha[k] = alpha[k]; g_w[k] = w[k]; g_q[k] = q[k]; }
}

__global__ void nprior_init_kernel(NPParams P)
{
    const int c = blockIdx.x;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    for (int j = threadIdx.x; j < P.p; j += blockDim.x) {
        const double z = P.alpha_init ? P.alpha_init[(int64_t)c * P.p + j] : rng_normal(key, NP_INIT_ALPHA, j, 0, 0);
        P.alpha[(int64_t)c * P.p + j] = 1 / sqrt(2 * log((double)P.p)) * z;      // :98
        P.w[(int64_t)c * P.p + j] = 0.0;
        P.beta[(int64_t)c * P.p + j] = 0.0;
    }
    if (threadIdx.x == 0) {
        const double z = P.sigma_init ? P.sigma_init[c] : rng_normal(key, NP_SIGMA_INIT, 0, 0, 0);
        P.hs[4 * c] = z * z;                                                   // :117
        P.hs[4 * c + 1] = norm_quantile_apprx_hd(P.cfg.p0_init);               // :119
        P.hs[4 * c + 2] = P.cfg.p0_init;
        P.hs[4 * c + 3] = P.yy;
    }
}

__global__ void pack_nonzero_kernel(const double* __restrict__ dense, int64_t rows, int64_t p, int words,
                                    int64_t N_pad, uint32_t* __restrict__ bits)
{
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= rows * words) return;
    const int64_t r = gw / words;
    const int w = (int)(gw % words);
    const int64_t col = 32LL * w + (threadIdx.x & 31);
    const bool nz = col < p && dense[r * p + col] != 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0) bits[(int64_t)w * N_pad + r] = m;
}

struct DBuf {
    void* ptr = nullptr; size_t bytes = 0;
    int alloc(size_t b) {
        bytes = b;
        if (!b) return BLIP_OK;
        cudaError_t e = cudaMalloc(&ptr, b);
        if (e != cudaSuccess) { ptr = nullptr; blip_set_error("cudaMalloc(%zu bytes) failed: %s", b, cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
        return BLIP_OK;
    }
    int upload(const void* host, size_t b, cudaStream_t st) {
        BLIP_TRY(alloc(b));
        CUDA_TRY(cudaMemcpyAsync(ptr, host, b, cudaMemcpyHostToDevice, st));
        return BLIP_OK;
    }
    ~DBuf() { if (ptr) cudaFree(ptr); }
    template <class T> T* as() { return reinterpret_cast<T*>(ptr); }
};

} // namespace

extern "C" int blipgpu_sample_nprior(blipgpu_ctx* ctx, blipgpu_design* design, const blipgpu_nprior_config* cfg,
                                     const double* y, int64_t chains, int64_t chain_offset, int64_t n_keep,
                                     int64_t burn, uint64_t seed, const blipgpu_nprior_streams* st_in,
                                     blipgpu_nprior** out)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!design || !cfg || !y || !out || chains <= 0 || n_keep <= 0 || burn < 0) { blip_set_error("sample_nprior: bad arguments"); return BLIP_ERR_ARG; }
    if (design->kind != BLIPGPU_DESIGN_DENSE) { blip_set_error("sample_nprior: needs a dense design"); return BLIP_ERR_ARG; }
    if (!(cfg->tauw2 > 0)) { blip_set_error("sample_nprior: tauw2 must be positive"); return BLIP_ERR_ARG; }
    BLIP_TRY(blipgpu_design_build_gram(design));
    const int64_t n = design->n, p = design->p, p2 = round_up(p, 2), n_total = n_keep + burn;
    cudaStream_t st = ctx->stream;

    blipgpu_nprior* res = new (std::nothrow) blipgpu_nprior();
    if (!res) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    res->ctx = ctx; res->rows = chains * n_keep; res->p = p; res->chains = chains; res->n_keep = n_keep;
    res->dense[0] = res->dense[1] = res->dense[2] = nullptr; res->scal = nullptr; res->sweep_ms = 0.0;
    struct Guard { blipgpu_nprior* r; ~Guard() { if (r) blipgpu_nprior_free(r); } } guard{ res };
    for (int i = 0; i < 3; ++i) {
        cudaError_t e = cudaMalloc(&res->dense[i], sizeof(double) * res->rows * p);
        if (e != cudaSuccess) {
            blip_set_error("sample_nprior: the dense alphas/ws/betas outputs need 3 x %.2f GB of device memory: %s",
                           8.0 * res->rows * p / 1e9, cudaGetErrorString(e));
            return BLIP_ERR_NOMEM;
        }
    }
    CUDA_TRY(cudaMalloc(&res->scal, sizeof(double) * 3 * res->rows));

    DBuf d_alpha, d_w, d_beta, d_q, d_uz, d_hs, d_chol, d_vec, d_act, d_err, d_q0, d_y, d_perm;
    DBuf i_u, i_wz, i_ai, i_wf, i_si, i_ja, i_jc, i_dz, i_ig, i_p0;
    BLIP_TRY(d_alpha.alloc(sizeof(double) * chains * p)); BLIP_TRY(d_w.alloc(sizeof(double) * chains * p));
    BLIP_TRY(d_beta.alloc(sizeof(double) * chains * p)); BLIP_TRY(d_q.alloc(sizeof(double) * chains * p2));
    BLIP_TRY(d_uz.alloc(sizeof(NPDraws) * chains * p));
    BLIP_TRY(d_hs.alloc(sizeof(double) * chains * 4));
    // active-set workspace of the joint W draw: starts small and is doubled (with a replay of the
    // launch, see below) when a chain's active set outgrows it; the limit is what half of the
    // free memory can hold
    // (BLIPGPU_NPRIOR_WORKSPACE_MB overrides the budget; the tests use it to force waves).  When the
    // workspace of all chains does not fit, the chains of a launch run in waves of `wave` chains.
    double budget = 0.0;
    {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        budget = 0.5 * (double)free_b;
        if (const char* e = getenv("BLIPGPU_NPRIOR_WORKSPACE_MB")) budget = std::max(1e-3, atof(e)) * 1048576.0;
    }
    const int64_t amax_limit = std::min<int64_t>(p, std::max<int64_t>(NP_AMAX_MIN, (int64_t)std::sqrt(budget / 8.0)));
    int64_t NP_AMAX = std::min<int64_t>(amax_limit, 256);
    int64_t wave = chains;
    auto alloc_workspace = [&]() -> int {
        if (d_chol.ptr) { cudaFree(d_chol.ptr); d_chol.ptr = nullptr; }
        if (d_vec.ptr) { cudaFree(d_vec.ptr); d_vec.ptr = nullptr; }
        if (d_act.ptr) { cudaFree(d_act.ptr); d_act.ptr = nullptr; }
        wave = std::max<int64_t>(1, std::min<int64_t>(chains, (int64_t)(budget / (8.0 * (double)NP_AMAX * (double)NP_AMAX))));
        BLIP_TRY(d_chol.alloc(sizeof(double) * wave * NP_AMAX * NP_AMAX));
        BLIP_TRY(d_vec.alloc(sizeof(double) * wave * 3 * NP_AMAX)); BLIP_TRY(d_act.alloc(sizeof(int) * wave * NP_AMAX));
        return BLIP_OK;
    };
    BLIP_TRY(alloc_workspace());
    DBuf s_alpha, s_w, s_beta, s_q, s_hs;            // state snapshot for the replay
    BLIP_TRY(s_alpha.alloc(d_alpha.bytes)); BLIP_TRY(s_w.alloc(d_w.bytes)); BLIP_TRY(s_beta.alloc(d_beta.bytes));
    BLIP_TRY(s_q.alloc(d_q.bytes)); BLIP_TRY(s_hs.alloc(d_hs.bytes));
    BLIP_TRY(d_err.alloc(2 * sizeof(int))); BLIP_TRY(d_q0.alloc(sizeof(double) * p));
    BLIP_TRY(d_y.upload(y, sizeof(double) * n, st));
    CUDA_TRY(cudaMemsetAsync(d_err.ptr, 0, 2 * sizeof(int), st));
    BLIP_TRY(blip_xty(st, design, d_y.as<double>(), d_q0.as<double>()));
    double yy = 0.0;
    for (int64_t i = 0; i < n; ++i) yy += y[i] * y[i];

    NPParams P;
    memset(&P, 0, sizeof(P));
    P.G = design->G; P.ldg = design->ldg; P.Xl2 = design->Xl2; P.q0 = d_q0.as<double>(); P.yy = yy;
    P.n = (int)n; P.p = (int)p; P.p2 = (int)p2; P.cfg = *cfg;
    // chain state in shared memory when 4 p doubles fit next to the kernel's static shared memory
    size_t dyn_smem = 0;
    {
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, nprior_kernel));
        const size_t want = sizeof(double) * 4 * (size_t)p2;
        const size_t room = (size_t)ctx->smem_optin > fa.sharedSizeBytes + 1024 ? (size_t)ctx->smem_optin - fa.sharedSizeBytes - 1024 : 0;
        if (want <= room) {
            dyn_smem = want;
            CUDA_TRY(cudaFuncSetAttribute(nprior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        }
        P.state_smem = dyn_smem ? 1 : 0;
    }
    P.min_alpha0 = norm_quantile_apprx_hd(cfg->min_p0);
    P.sigma_a = n / 2.0 + cfg->sigma_a0 + (cfg->sigma_prior_type != 0 ? p / 2.0 : 0.0);
    P.alpha = d_alpha.as<double>(); P.w = d_w.as<double>(); P.beta = d_beta.as<double>(); P.q = d_q.as<double>();
    P.uz = d_uz.as<NPDraws>();
    P.amax = (int)NP_AMAX;
    P.hs = d_hs.as<double>(); P.chol = d_chol.as<double>(); P.vec = d_vec.as<double>(); P.act = d_act.as<int>();
    P.err = d_err.as<int>();
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); P.chain_offset = (uint32_t)chain_offset;
    P.n_total = (int)n_total; P.burn = (int)burn; P.n_keep = (int)n_keep;
    for (int i = 0; i < 3; ++i) P.out[i] = res->dense[i];
    P.scal = res->scal; P.rows = res->rows;
    if (st_in) {
        const size_t big = sizeof(double) * chains * n_total * p, vecN = sizeof(double) * chains * n_total;
        if (st_in->u) { BLIP_TRY(i_u.upload(st_in->u, big, st)); P.u = i_u.as<double>(); }
        if (st_in->wz) { BLIP_TRY(i_wz.upload(st_in->wz, big, st)); P.wz = i_wz.as<double>(); }
        if (st_in->alpha_init) { BLIP_TRY(i_ai.upload(st_in->alpha_init, sizeof(double) * chains * p, st)); P.alpha_init = i_ai.as<double>(); }
        if (st_in->w_fresh) { BLIP_TRY(i_wf.upload(st_in->w_fresh, big, st)); P.w_fresh = i_wf.as<double>(); }
        if (st_in->sigma_init) { BLIP_TRY(i_si.upload(st_in->sigma_init, sizeof(double) * chains, st)); P.sigma_init = i_si.as<double>(); }
        if (st_in->joint_all) { BLIP_TRY(i_ja.upload(st_in->joint_all, big, st)); P.joint_all = i_ja.as<double>(); }
        if (st_in->joint_act) { BLIP_TRY(i_jc.upload(st_in->joint_act, big, st)); P.joint_act = i_jc.as<double>(); }
        if (st_in->delta_z) { BLIP_TRY(i_dz.upload(st_in->delta_z, vecN, st)); P.delta_z = i_dz.as<double>(); }
        if (st_in->invgamma) { BLIP_TRY(i_ig.upload(st_in->invgamma, vecN, st)); P.invgamma = i_ig.as<double>(); }
        if (st_in->p0_draw) { BLIP_TRY(i_p0.upload(st_in->p0_draw, vecN, st)); P.p0_draw = i_p0.as<double>(); }
    }
    nprior_init_kernel<<<(unsigned)chains, 256, 0, st>>>(P);
    CUDA_TRY(cudaGetLastError());

    // permutations: chunks of S sweeps generated on the same stream
    const bool inj_perm = st_in && st_in->perm;
    const bool perm16 = !inj_perm && p <= 65535;
    int S = (int)std::min<int64_t>(n_total, std::max<int64_t>(1, (int64_t)(5e8 / ((perm16 ? 2.0 : 4.0) * chains * p))));
    S = std::min(S, 256);
    BLIP_TRY(d_perm.alloc((perm16 ? 2 : 4) * (size_t)chains * S * p));
    for (int64_t sweep0 = 0; sweep0 < n_total; sweep0 += S) {
        const int Sc = (int)std::min<int64_t>(S, n_total - sweep0);
        if (inj_perm) {
            CUDA_TRY(cudaMemcpy2DAsync(d_perm.ptr, sizeof(int32_t) * Sc * p, st_in->perm + sweep0 * p,
                                       sizeof(int32_t) * n_total * p, sizeof(int32_t) * Sc * p, (size_t)chains,
                                       cudaMemcpyHostToDevice, st));
        } else {
            BLIP_TRY(blip_fill_perms(st, d_perm.ptr, perm16 ? 1 : 0, chains, Sc, p, sweep0, P.k0, P.k1, P.chain_offset));
            ctx->kernel_launches += 1;
        }
        P.perm = d_perm.ptr; P.perm16 = perm16 ? 1 : 0; P.perm_S = Sc; P.perm_s0 = 0;
        P.sweep0 = (int)sweep0; P.S = Sc;
        auto copy_state = [&](bool save) -> int {
            DBuf* live[5] = { &d_alpha, &d_w, &d_beta, &d_q, &d_hs };
            DBuf* snap[5] = { &s_alpha, &s_w, &s_beta, &s_q, &s_hs };
            for (int i = 0; i < 5; ++i)
                CUDA_TRY(cudaMemcpyAsync(save ? snap[i]->ptr : live[i]->ptr, save ? live[i]->ptr : snap[i]->ptr,
                                         live[i]->bytes, cudaMemcpyDeviceToDevice, st));
            return BLIP_OK;
        };
        BLIP_TRY(copy_state(true));
        for (;;) {
            P.amax = (int)NP_AMAX; P.chol = d_chol.as<double>(); P.vec = d_vec.as<double>(); P.act = d_act.as<int>();
            ScopedTimer timer(ctx, true);
            for (int64_t base = 0; base < chains; base += wave) {
                P.chain_base = (int)base;
                nprior_kernel<<<(unsigned)std::min<int64_t>(wave, chains - base), NP_BLOCK, dyn_smem, st>>>(P);
                CUDA_TRY(cudaGetLastError());
            }
            res->sweep_ms += timer.stop(1);
            int errv[2] = { 0, 0 };
            CUDA_TRY(cudaMemcpyAsync(errv, d_err.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const int err = errv[0];
            if (err == 2) { blip_set_error("sample_nprior: Cholesky of the active-set precision failed"); return BLIP_ERR_NUMERIC; }
            if (err == 0) break;
            // a chain's active set outgrew the workspace: enlarge it and replay the launch from the
            // snapshot (keyed draws make the replay identical)
            if (NP_AMAX >= amax_limit) {
                blip_set_error("sample_nprior: more than %lld active coordinates in the joint W draw (one chain's workspace exceeds half of the free memory)", (long long)NP_AMAX);
                return BLIP_ERR_UNSUPPORTED;
            }
            NP_AMAX = std::min<int64_t>(amax_limit, std::max<int64_t>(2 * NP_AMAX, (int64_t)errv[1] + errv[1] / 4 + 32));
            BLIP_TRY(alloc_workspace());
            BLIP_TRY(copy_state(false));
            CUDA_TRY(cudaMemsetAsync(d_err.ptr, 0, 2 * sizeof(int), st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    guard.r = nullptr;
    *out = res;
    return BLIP_OK;
}

extern "C" int blipgpu_nprior_get(blipgpu_nprior* r, int32_t which, double* out)
{
    if (!r || !out || which < 0 || which > 5) { blip_set_error("nprior_get: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(r->ctx));
    if (which < 3) {
        BLIP_TRY(blip_d2h(r->ctx, out, r->dense[which], sizeof(double) * r->rows * r->p));
        r->ctx->d2h_bytes += (int64_t)sizeof(double) * r->rows * r->p;
    } else {
        CUDA_TRY(cudaMemcpy(out, r->scal + (which - 3) * r->rows, sizeof(double) * r->rows, cudaMemcpyDeviceToHost));
        r->ctx->d2h_bytes += (int64_t)sizeof(double) * r->rows;
    }
    return BLIP_OK;
}

extern "C" int blipgpu_nprior_info(blipgpu_nprior* r, int64_t* rows, int64_t* p, double* sweep_ms)
{
    if (!r) { blip_set_error("nprior_info: NULL handle"); return BLIP_ERR_ARG; }
    if (rows) *rows = r->rows;
    if (p) *p = r->p;
    if (sweep_ms) *sweep_ms = r->sweep_ms;
    return BLIP_OK;
}

extern "C" int blipgpu_nprior_bits(blipgpu_nprior* r, blipgpu_bits** out)
{
    if (!r || !out) { blip_set_error("nprior_bits: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(r->ctx));
    blipgpu_bits* b = nullptr;
    BLIP_TRY(blip_bits_alloc(r->ctx, r->rows, r->p, &b));
    const int64_t nwarps = r->rows * b->words;
    pack_nonzero_kernel<<<(unsigned)((nwarps + 7) / 8), 256, 0, r->ctx->stream>>>(r->dense[2], r->rows, r->p, (int)b->words, b->N_pad, b->bits);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->ctx->stream);
    if (e != cudaSuccess) { blipgpu_bits_free(b); blip_set_error("nprior_bits: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    *out = b;
    return BLIP_OK;
}

extern "C" int blipgpu_nprior_free(blipgpu_nprior* r)
{
    if (!r) return BLIP_OK;
    cudaSetDevice(r->ctx->device);
    for (int i = 0; i < 3; ++i) cudaFree(r->dense[i]);
    cudaFree(r->scal);
    delete r;
    return BLIP_OK;
}
