// gibbs_gram_stream.cuh -- gram-form sweep for few chains as a CHECKED COLUMN STREAM on a 2-CTA cluster.
//
// The other gram kernels walk the visiting order with a speculation window and pay, per change, a
// round of gathered Gram entries (one scattered 8-byte load per window position: a fully divergent
// warp load occupies the L1 tag stage for 32 cycles) on the chain's critical path.  In the stationary
// regime that is the wrong unit of work: of the K ~ 61 changes of a sweep at C2, ~ 55 are redraws of
// coordinates that were already active when the sweep started, their positions in the visiting order
// are known up front, and what such a redraw needs is q at the ACTIVE coordinates only.  So the sweep
// is split, per chain, over the two SMs of a thread-block cluster:
//
//   control CTA (rank 0)
//     R  resolver (warp 15)    keeps q at up to ST_NA tracked active coordinates in registers (two per
//                              lane) and the Gram sub-matrix G[A][A] in shared memory, and decides the
//                              tracked redraws one after the other in visiting order -- decision, slab
//                              draw, delta, ||r||^2 -- without touching global memory, speculating that
//                              no other coordinate changes in between.  It has an SM's load/store
//                              pipeline to itself: on the streamers' SM every shared-memory access or
//                              shuffle of the chain queued behind their 16-byte loads (measured: 2200
//                              cycles per redraw instead of ~400);
//     E  warps 8-15            sweep start (decision coefficients by coordinate, active list in visiting
//                              order, tracked set) and the sweep epilogue (hyper-parameters, recording);
//     P  precompute (warps 0-7) the next sweep's state-independent tables, by COORDINATE: position of
//                              every coordinate in the visiting order, logit of its spike uniform;
//   stream CTA (rank 1)
//     S  streamers (16 warps)  own q (twice, in shared memory), verify the resolver's speculation and
//                              apply the changes in one pass over q: a pass takes the next <= ST_MAXC
//                              decided changes, streams their Gram columns with coalesced 16-byte loads,
//                              q_dst = q_src - sum_i delta_i G[:, j_i] in change order, and checks every
//                              coordinate whose visit falls in the pass's position range against q as
//                              it stood at that visit (the value after exactly the changes that precede
//                              the visit): coordinate order, no gathers.  The coordinates to check go
//                              first; if one of them changes the pass is dropped before the bulk of the
//                              stream, and the source buffer is intact either way.
//
// If a check fails at position f (an activation, ~ 5 per sweep, or an active coordinate beyond the
// tracked ones), the changes before f are final, the owner of f computes its change exactly like the
// classic kernel's owner thread, R rewinds to its snapshot after the last final change, applies the
// new change (one gathered Gram entry per tracked coordinate) and decides the remaining redraws again.
// The two CTAs exchange a pass descriptor and a result record through distributed shared memory,
// a few words per pass; each side polls its own shared memory only.
//
// Every decision is the exact one (the guarded fast test of gibbs_gram_kernel or the reference formula),
// every coefficient is drawn by the same expression, and every element of q sees the same fma
// sequence in change order as in gibbs_gram_kernel: results are bit-identical to it
// (tests/test_stream_gpu.py).  The work counter n_windows counts passes here.
//
// Reference semantics: /root/reference/pyblip/linear/_linear.pyx:126-229 (unchanged).
#pragma once
#include "gibbs_gram_batch.cuh"

constexpr int ST_THREADS = 512;
constexpr int ST_NP = 256;                 // control CTA: precompute, warps 0-7 (lowest arbitration priority)
constexpr int ST_E0 = 256;                 // control CTA: sweep start / epilogue group, warps 8-15
constexpr int ST_R0 = 480;                 // control CTA: resolver, warp 15
constexpr int ST_NS = 512;                 // stream CTA: every thread
#ifndef BLIP_ST_MAXC
#define BLIP_ST_MAXC 8
#endif
constexpr int ST_MAXC = BLIP_ST_MAXC;      // changes applied per pass
constexpr int ST_NA = 64;                  // tracked active coordinates (two per resolver lane)
constexpr int ST_RING = 32;                // decided changes kept (ring), with the resolver's snapshot after each
constexpr int ST_LOOK = 16;                // decided-but-uncommitted changes the resolver may run ahead
constexpr int ST_LIST = 256;               // active coordinates listed at sweep start
constexpr int ST_CHUNK = 8;                // coordinate pairs a streamer looks up at once
#ifndef BLIP_ST_IDLE
#define BLIP_ST_IDLE 3
#endif
constexpr int ST_IDLE = BLIP_ST_IDLE;      // redraws decided with the streamers idle before a short list is handed over
constexpr int ST_NONE = 0x7fffffff;
enum { ST_CMD_PASS = 1, ST_CMD_SWEEP_BEGIN = 2, ST_CMD_REFRESH = 3, ST_CMD_QUIT = 4 };
typedef SyncNamed<1, GB_BLOCK, ST_E0> SyncE;
typedef SyncNamed<4, ST_NP> SyncPre;

#ifdef BLIP_PHASE_CLOCKS
#define STP_DECL long long sp_t[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0}; long long sp_last = clock64();
#define STP(i) do { if ((threadIdx.x & 31) == 0) { long long _n = clock64(); sp_t[i] += _n - sp_last; sp_last = _n; } } while (0)
#define STP_CNT(i, v) do { sp_t[i] += (v); } while (0)
#define STP_ARGS , long long* sp_t, long long& sp_last
#define STP_PASS , sp_t, sp_last
#else
#define STP_DECL
#define STP(i) do {} while (0)
#define STP_CNT(i, v) do {} while (0)
#define STP_ARGS
#define STP_PASS
#endif

struct StChange { int pos, j; double delta, bnew, rr; };

// lives in the stream CTA; the resolver writes it through distributed shared memory
struct StPass {
    int cmd_seq, cmd_type;                // R -> S
    int s_cmd;                            // S internal
    int n, lo, hi, src;                   // pass: changes, check range (lo, hi], source buffer; SWEEP_BEGIN: lo = sweep
    int j[ST_MAXC], pos[ST_MAXC];
    double d[ST_MAXC];
    int s_fail;                           // first changing position found so far
};

// lives in the control CTA; the streamers write the result fields through distributed shared memory
struct StCtl {
    SweepShared sh;
    int pre_ready, d_sweep;               // P -> E: sweeps < pre_ready have their tables; E -> P: sweep being run
    int done_seq, r_fail;                 // S -> R
    int f_j; double f_t, f_xl2, f_delta, f_bnew;
    int buf;                              // q buffer that holds every committed change
    int act_n, T;
    int act_pos[ST_LIST], act_j[ST_LIST];
    int srt_pos[ST_NA], srt_j[ST_NA];
    float srt_fA[ST_NA], srt_fc1[ST_NA], srt_fL[ST_NA];
    double srt_bold[ST_NA], srt_xl2[ST_NA], srt_zpx[ST_NA], srt_zpy[ST_NA];
    StChange ring[ST_RING];
    double rr_init;
    double scratch[40];
};

__host__ __device__ inline size_t st_smem_bytes(int64_t p2, int64_t words)
{
    return sizeof(double) * (2 * (size_t)p2 + (size_t)ST_NA * ST_NA + (size_t)(ST_RING + 1) * ST_NA) + 2 * sizeof(uint32_t) * (size_t)words;
}

__device__ __forceinline__ void st_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// =================================================================================================
// P: tables of sweep pre_next by coordinate -- inv[j] = position of j, Lc[j] = logit of its spike uniform
// =================================================================================================
__device__ __forceinline__ void st_precompute_role(const SweepParams& P, int c, StCtl& ctl)
{
    const int ptid = threadIdx.x;
    const int p = P.p, p2 = P.p2;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    const bool keyed_perm = P.perm == nullptr;
    for (int pre = 0; pre < P.S; ++pre) {
        if (ptid == 0) {
            while (pre > ld_volatile_int(&ctl.d_sweep) + 1) __nanosleep(200);
            __threadfence_block();
        }
        SyncPre::sync();
        const int sweep = P.sweep0 + pre;
        int* inv = P.scrI + (int64_t)(pre & 1) * P.n_chains * p2 + (int64_t)c * p2;
        float* Lc = P.scrL + (int64_t)(pre & 1) * P.scr_stride + (int64_t)c * p;
        PermKey pk;
        if (keyed_perm) pk = perm_key(P.k0, P.k1, key.chain, (uint32_t)sweep, (uint32_t)p);
        const PermView pv = perm_view(P, c, pre);
        const double* u_s = P.u ? P.u + ((int64_t)c * P.S + pre) * p : nullptr;
        for (int pos = ptid; pos < p; pos += ST_NP) {
            const int j = keyed_perm ? (int)perm_at(pk, (uint32_t)pos, (uint32_t)p) : pv[pos];
            const double u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
            const bool ex = !(u > 1e-4 && u < 1.0 - 1e-4);
            inv[j] = pos;
            Lc[j] = ex ? __int_as_float(0x7fc00000) : __logf((float)u) - __logf((float)(1.0 - u));
        }
        if (ptid == 0 && p2 > p) inv[p] = -1;
        __threadfence();
        SyncPre::sync();
        if (ptid == 0) { __threadfence_block(); st_volatile_int(&ctl.pre_ready, pre + 1); }
    }
}

// resolver warp: hand a command to the streamers / wait for its acknowledgement
__device__ __forceinline__ void st_issue(StPass* ps, int type, int& seq)
{
    st_fence();                                                   // every lane's descriptor stores
    __syncwarp();
    ++seq;
    if ((threadIdx.x & 31) == 0) { ps->cmd_type = type; st_fence(); st_volatile_int(&ps->cmd_seq, seq); }
}
__device__ __forceinline__ bool st_done(const StCtl& ctl, int seq)
{
    int d = ld_volatile_int(&ctl.done_seq);
    d = __shfl_sync(0xffffffffu, d, 0);
    if (d != seq) return false;
    st_fence();
    return true;
}

// =================================================================================================
// E (warps 8-15 of the control CTA): sweep start -- decision coefficients by coordinate, the active
// coordinates in visiting order, the tracked set with its slab draws and its Gram sub-matrix
// =================================================================================================
template <int DESIGN>
__device__ __forceinline__ void st_sweep_start(const SweepParams& P, int c, int s, StCtl& ctl, StPass* ps, const double* q_r, double* GA,
                                               uint32_t* mask_s, uint32_t* trk_s, int& seq)
{
    const int tid = SyncE::tid();
    const bool is_r = threadIdx.x >= ST_R0;
    const int p = P.p, p2 = P.p2;
    const int sweep = P.sweep0 + s;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    SweepShared& sh = ctl.sh;
    double* beta_g = P.beta + (int64_t)c * p;
    if (tid == 0) {
        st_volatile_int(&ctl.d_sweep, s);
        while (ld_volatile_int(&ctl.pre_ready) <= s) __nanosleep(100);
        __threadfence_block();
        __threadfence();
    }
    // ---- periodic rebuild of q and ||r||^2 from q0 = X^T y (as in the other gram kernels: the streamers redo q
    //      with the same fma per element, ||r||^2 is summed here by the same 256-thread tree)
    if (P.gram_refresh > 0 && sweep > 0 && sweep % P.gram_refresh == 0) {
        if (is_r) {
            if ((threadIdx.x & 31) == 0) ps->src = ctl.buf;
            st_issue(ps, ST_CMD_REFRESH, seq);
            while (!st_done(ctl, seq)) __nanosleep(100);
        }
        SyncE::sync();
        const double* q = q_r + (size_t)ctl.buf * p2;
        double acc = 0.0;
        for (int w = tid; w < P.words; w += GB_BLOCK) {
            uint32_t m = mask_s[w];
            while (m) {
                int j = 32 * w + __ffs(m) - 1;
                m &= m - 1;
                acc = fma(__ldcg(beta_g + j), P.q0[j] + q[j], acc);
            }
        }
        acc = block_sum<SyncE>(acc, ctl.scratch);
        if (tid == 0) sh.rr = P.yy - acc;
    }
    SyncE::sync();
    if (tid == 0) { sh.logodds = log(sh.p0) - log(1 - sh.p0); ctl.act_n = 0; }
    for (int w = tid; w < P.words; w += GB_BLOCK) trk_s[w] = 0u;
    SyncE::sync();
    const Hyper h = { sh.sigma2, sh.tau2, sh.logodds };
    const float f_logodds = (float)h.logodds, f_tau2 = (float)h.tau2, f_sigma2 = (float)h.sigma2;
    const float f_ts = (float)(h.tau2 / h.sigma2);
    const int* inv = P.scrI + (int64_t)(s & 1) * P.n_chains * p2 + (int64_t)c * p2;
    const float* Lc = P.scrL + (int64_t)(s & 1) * P.scr_stride + (int64_t)c * p;
    float4* aux = reinterpret_cast<float4*>(P.scrZ) + (int64_t)c * p;
    const double* z_s = P.z ? P.z + ((int64_t)c * P.S + s) * p : nullptr;

    // coefficients of the guarded fast decision  E = A - c1 q^2 >= L, by coordinate: (A - L, c1, guard base)
    for (int k = tid; k < p; k += GB_BLOCK) {
        const float xl2f = (float)P.Xl2[k];
        const float L = __ldcg(Lc + k);
        const float fA = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
        const float fc1 = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
        const bool ok = (L == L) && fmaxf(fA, L) < 20.0f;
        __stcg(aux + k, make_float4(ok ? fA - L : __int_as_float(0x7fc00000), fc1, 1e-4f * (1.0f + fabsf(fA) + fabsf(L)), 0.f));
    }
    for (int w = tid; w < P.words; w += GB_BLOCK) {
        uint32_t m = mask_s[w];
        while (m) {
            const int j = 32 * w + __ffs(m) - 1;
            m &= m - 1;
            const int slot = atomicAdd(&ctl.act_n, 1);
            if (slot < ST_LIST) { ctl.act_j[slot] = j; ctl.act_pos[slot] = __ldcg(inv + j); }
        }
    }
    SyncE::sync();
    // more active coordinates than the list holds (early burn-in): nothing is tracked, every active visit
    // is found by the streamers
    const int n_list = ctl.act_n <= ST_LIST ? ctl.act_n : 0;
    const int T = min(n_list, ST_NA);
    if (tid < n_list) {
        const int pos = ctl.act_pos[tid], j = ctl.act_j[tid];
        int rank = 0;
        for (int k = 0; k < n_list; ++k) rank += ctl.act_pos[k] < pos;
        if (rank < ST_NA) {
            const double zn = z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0);
            const double xl2 = P.Xl2[j];
            const double pv = 1.0 / (1.0 / h.tau2 + xl2 / h.sigma2);
            ctl.srt_pos[rank] = pos; ctl.srt_j[rank] = j;
            ctl.srt_bold[rank] = __ldcg(beta_g + j); ctl.srt_xl2[rank] = xl2;
            ctl.srt_zpx[rank] = sqrt(pv) * zn; ctl.srt_zpy[rank] = pv;
            ctl.srt_fL[rank] = __ldcg(Lc + j);
            const float xl2f = (float)xl2;
            ctl.srt_fA[rank] = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
            ctl.srt_fc1[rank] = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
            atomicOr(&trk_s[j >> 5], 1u << (j & 31));
        }
    }
    if (tid == 0) ctl.T = T;
    SyncE::sync();
    for (int idx = tid; idx < T * T; idx += GB_BLOCK) {
        const int a = idx / T, t = idx - a * T;
        GA[a * ST_NA + t] = gram_entry<DESIGN>(P, ctl.srt_j[a], ctl.srt_j[t]);
    }
    SyncE::sync();
    // the streamers take their copies of the active set / tracked set and the hyper-parameters
    if (is_r) {
        if ((threadIdx.x & 31) == 0) ps->lo = s;
        st_issue(ps, ST_CMD_SWEEP_BEGIN, seq);
        while (!st_done(ctl, seq)) __nanosleep(100);
    }
}

// =================================================================================================
// S (the stream CTA): commands from the resolver until it quits
// =================================================================================================
template <int DESIGN>
__device__ __forceinline__ void st_streamer_cta(const SweepParams& P, int c, StPass& pass, StCtl* rc, double* qbuf,
                                                uint32_t* mask_s, uint32_t* trk_s, const uint32_t* mask_r, const uint32_t* trk_r)
{
    const int tid = threadIdx.x;
    const int p = P.p, p2 = P.p2;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    double* beta_g = P.beta + (int64_t)c * p;
    const float4* aux = reinterpret_cast<const float4*>(P.scrZ) + (int64_t)c * p;
    Hyper h = { 1.0, 1.0, 0.0 };
    int s = 0, sweep = P.sweep0;
    const int* inv = P.scrI;
    const double* u_s = nullptr;
    const double* z_s = nullptr;
    STP_DECL
    for (int seq = 1;; ++seq) {
        if (tid == 0) {
            while (ld_volatile_int(&pass.cmd_seq) < seq) __nanosleep(20);
            st_fence();
            pass.s_cmd = pass.cmd_type;
        }
        __syncthreads();
        const int cmd = pass.s_cmd;
        STP(5);
        if (cmd == ST_CMD_QUIT) {
            const double* q = qbuf + (size_t)pass.src * p2;
            double* q_g = P.q + (int64_t)c * p2;
            for (int k = tid; k < p2; k += ST_NS) q_g[k] = q[k];
            break;
        }
        if (cmd == ST_CMD_SWEEP_BEGIN) {
            s = pass.lo; sweep = P.sweep0 + s;
            inv = P.scrI + (int64_t)(s & 1) * P.n_chains * p2 + (int64_t)c * p2;
            u_s = P.u ? P.u + ((int64_t)c * P.S + s) * p : nullptr;
            z_s = P.z ? P.z + ((int64_t)c * P.S + s) * p : nullptr;
            h.sigma2 = rc->sh.sigma2; h.tau2 = rc->sh.tau2; h.logodds = rc->sh.logodds;
            for (int w = tid; w < P.words; w += ST_NS) { mask_s[w] = mask_r[w]; trk_s[w] = trk_r[w]; }
            __syncthreads();
            if (tid == 0) { st_fence(); st_volatile_int(&rc->done_seq, seq); }
            STP(8);
            continue;
        }
        if (cmd == ST_CMD_REFRESH) {
            // q = q0 - sum over the active coordinates (ascending) of beta_j G[:, j]
            double* q = qbuf + (size_t)pass.src * p2;
            for (int w = tid; w < P.words; w += ST_NS) mask_s[w] = mask_r[w];
            for (int k = tid; k < p2; k += ST_NS) q[k] = k < p ? P.q0[k] : 0.0;
            __syncthreads();
            for (int w = 0; w < P.words; ++w) {
                uint32_t m = mask_s[w];
                while (m) {
                    const int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    const double bj = __ldcg(beta_g + j);
                    for (int k = 2 * tid; k < p2; k += 2 * ST_NS) {
                        const double2 g = gram_pair<DESIGN>(P, j, k);
                        double2 qq = *reinterpret_cast<double2*>(q + k);
                        qq.x = fma(-bj, g.x, qq.x); qq.y = fma(-bj, g.y, qq.y);
                        *reinterpret_cast<double2*>(q + k) = qq;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { st_fence(); st_volatile_int(&rc->done_seq, seq); }
            continue;
        }
        // ---- a pass
        const int n = pass.n, lo = pass.lo;
        const unsigned span = (unsigned)(pass.hi - lo);
        const double* src = qbuf + (size_t)pass.src * p2;
        double* dst = qbuf + (size_t)(pass.src ^ 1) * p2;
        int cj[ST_MAXC];
        double cd[ST_MAXC];
#pragma unroll
        for (int i = 0; i < ST_MAXC; ++i) { cj[i] = i < n ? pass.j[i] : 0; cd[i] = i < n ? pass.d[i] : 0.0; }
        unsigned done = 0u;                                        // own pairs already streamed (bit = iteration)
        int cand_pos = ST_NONE, cand_k = 0;
        double cand_t = 0.0;
        bool cand_act = false;
        if (span > 0u) {
            // ---- the pairs that hold a coordinate to check go first
            int cp[ST_MAXC];
#pragma unroll
            for (int i = 0; i < ST_MAXC; ++i) cp[i] = i < n ? pass.pos[i] : ST_NONE;
            auto check = [&](int k, int pos, double t, const float4& a) {
                const bool act = (mask_s[k >> 5] >> (k & 31)) & 1u;
                bool fail;
                if (act) {
                    if ((trk_s[k >> 5] >> (k & 31)) & 1u) return;      // the resolver's
                    fail = true;                                       // an active coordinate nobody tracks: its redraw
                } else {
                    const double x2c = (double)a.y * (t * t);
                    const double margin = (double)a.x - x2c;
                    const double guard = (double)a.z + 1e-4 * x2c;
                    if (fabs(margin) > guard) fail = margin < 0.0;
                    else {
                        const double c_u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                        fail = !(c_u <= kappa_of(t, P.Xl2[k], h));
                    }
                }
                if (fail) {
                    atomicMin(&pass.s_fail, pos);
                    if (pos < cand_pos) { cand_pos = pos; cand_k = k; cand_t = t; cand_act = act; }
                }
            };
            for (int it0 = 0, kk0 = 2 * tid; kk0 < p2; it0 += ST_CHUNK, kk0 += 2 * ST_NS * ST_CHUNK) {
                // look up the positions of a chunk of own pairs at once; the (few) pairs in range are then processed
                // by ONE copy of the code below (the position is read again with the pair's Gram loads)
                unsigned mx = 0u, my = 0u;
                {
                    int2 ipv[ST_CHUNK];
#pragma unroll
                    for (int u = 0; u < ST_CHUNK; ++u) {
                        const int kk = kk0 + u * 2 * ST_NS;
                        ipv[u] = kk < p2 ? __ldcg(reinterpret_cast<const int2*>(inv + kk)) : make_int2(-1, -1);
                    }
#pragma unroll
                    for (int u = 0; u < ST_CHUNK; ++u) {
                        mx |= ((unsigned)(ipv[u].x - lo - 1) < span ? 1u : 0u) << u;
                        my |= ((unsigned)(ipv[u].y - lo - 1) < span ? 1u : 0u) << u;
                    }
                }
                unsigned need = mx | my;
                done |= need << it0;
#pragma unroll 1
                while (need) {
                    const int u = __ffs(need) - 1;
                    need &= need - 1;
                    const int kk = kk0 + u * 2 * ST_NS;
                    bool inx = (mx >> u) & 1u, iny = (my >> u) & 1u;
                    const int2 ip = __ldcg(reinterpret_cast<const int2*>(inv + kk));
                    const int fcur = ld_volatile_int(&pass.s_fail);
                    if (fcur != ST_NONE) {                             // the pass is lost; only earlier positions still matter
                        inx = inx && ip.x < fcur; iny = iny && ip.y < fcur;
                        if (!inx && !iny) continue;
                    }
                    double2 g[ST_MAXC];
#pragma unroll
                    for (int i = 0; i < ST_MAXC; ++i) g[i] = i < n ? gram_pair<DESIGN>(P, cj[i], kk) : make_double2(0.0, 0.0);
                    float4 ax = make_float4(0.f, 0.f, 0.f, 0.f), ay = ax;
                    if (inx) ax = __ldcg(aux + kk);
                    if (iny) ay = __ldcg(aux + kk + 1);
                    double2 qq = *reinterpret_cast<const double2*>(src + kk);
                    double tx = qq.x, ty = qq.y;                       // q as it stands when the coordinate is visited
#pragma unroll
                    for (int i = 0; i < ST_MAXC; ++i) {
                        if (i < n) {
                            qq.x = fma(-cd[i], g[i].x, qq.x); qq.y = fma(-cd[i], g[i].y, qq.y);
                            if (cp[i] < ip.x) tx = qq.x;
                            if (cp[i] < ip.y) ty = qq.y;
                        }
                    }
                    *reinterpret_cast<double2*>(dst + kk) = qq;
                    if (inx) check(kk, ip.x, tx, ax);
                    if (iny) check(kk + 1, ip.y, ty, ay);
                }
            }
            STP(6);
            __syncthreads();
            const int f = pass.s_fail;
            if (f != ST_NONE) {
                if (cand_pos == f) {
                    // the owner of the first changing position: its change, exactly as gibbs_gram_kernel's owner thread
                    const int k = cand_k;
                    const double xl2 = P.Xl2[k];
                    const double bold = cand_act ? __ldcg(beta_g + k) : 0.0;
                    const double XjTr = cand_act ? cand_t + bold * xl2 : cand_t;
                    bool spike = false;
                    if (cand_act) {
                        const double c_u = u_s ? u_s[f] : rng_uniform(key, KIND_SPIKE_U, f, sweep, 0);
                        spike = c_u <= kappa_of(XjTr, xl2, h);
                    }
                    double bnew = 0.0;
                    if (!spike) bnew = slab_draw(XjTr, xl2, h, z_s ? z_s[f] : rng_normal(key, KIND_SLAB_Z, f, sweep, 0));
                    rc->f_j = k; rc->f_t = cand_t; rc->f_xl2 = xl2; rc->f_delta = bnew - bold; rc->f_bnew = bnew;
                    st_fence();
                }
                __syncthreads();
                if (tid == 0) { rc->r_fail = f; st_fence(); st_volatile_int(&rc->done_seq, seq); }
                STP(9);
                continue;
            }
        }
        // ---- the bulk: a plain stream of the columns over the remaining pairs
        {
            int it = 0;
            for (int kk = 2 * tid; kk < p2; kk += 2 * ST_NS, ++it) {
                if ((done >> it) & 1u) continue;
                double2 g[ST_MAXC];
#pragma unroll
                for (int i = 0; i < ST_MAXC; ++i) g[i] = i < n ? gram_pair<DESIGN>(P, cj[i], kk) : make_double2(0.0, 0.0);
                double2 qq = *reinterpret_cast<const double2*>(src + kk);
#pragma unroll
                for (int i = 0; i < ST_MAXC; ++i)
                    if (i < n) { qq.x = fma(-cd[i], g[i].x, qq.x); qq.y = fma(-cd[i], g[i].y, qq.y); }
                *reinterpret_cast<double2*>(dst + kk) = qq;
            }
        }
        STP(7);
        __syncthreads();
        if (tid == 0) { rc->r_fail = ST_NONE; st_fence(); st_volatile_int(&rc->done_seq, seq); }
        STP(9);
    }
#ifdef BLIP_PHASE_CLOCKS
    if (c == 0 && tid == 0)
        printf("st streamer | wait-cmd %lld check-phase %lld bulk-phase %lld post %lld sweep-begin %lld\n", sp_t[5], sp_t[6], sp_t[7], sp_t[9], sp_t[8]);
#endif
}

// =================================================================================================
// R: the tracked redraws in visiting order, pass planning, rewinds
// =================================================================================================
template <int DESIGN>
__device__ __forceinline__ void st_resolver_role(const SweepParams& P, int c, int s, StCtl& ctl, StPass* ps, const double* q_r, const double* GA,
                                                 double* snap, uint32_t* mask_s, int& seq, long long& n_changes, long long& n_passes STP_ARGS)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int p = P.p, p2 = P.p2;
    const int sweep = P.sweep0 + s;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    const Hyper h = { ctl.sh.sigma2, ctl.sh.tau2, ctl.sh.logodds };
    const double* u_s = P.u ? P.u + ((int64_t)c * P.S + s) * p : nullptr;
    double* beta_g = P.beta + (int64_t)c * p;
    const int T = ctl.T;
    int buf = ctl.buf;
    const int mj0 = lane < T ? ctl.srt_j[lane] : -1, mj1 = lane + 32 < T ? ctl.srt_j[lane + 32] : -1;
    const int mp0 = lane < T ? ctl.srt_pos[lane] : ST_NONE, mp1 = lane + 32 < T ? ctl.srt_pos[lane + 32] : ST_NONE;
    double qa0 = mj0 >= 0 ? q_r[(size_t)buf * p2 + mj0] : 0.0, qa1 = mj1 >= 0 ? q_r[(size_t)buf * p2 + mj1] : 0.0;
    double rr = ctl.sh.rr;
    snap[ST_RING * ST_NA + lane] = qa0; snap[ST_RING * ST_NA + lane + 32] = qa1;
    if (lane == 0) ctl.rr_init = rr;
    __syncwarp();

    int V = -1, ia = 0, n_dec = 0, n_com = 0, n_fin = 0;     // warp-uniform
    bool in_flight = false;
    int fl_n = 0, fl_hi = 0, idle_resolved = 0;
    STP(0);
    // the decided changes [n_fin, b) become final: coefficient and active-set bit
    auto finalize = [&](int b) {
        for (int idx = n_fin + lane; idx < b; idx += 32) {
            const StChange& e = ctl.ring[idx & (ST_RING - 1)];
            __stcg(beta_g + e.j, e.bnew);
            if (e.bnew != 0) atomicOr(&mask_s[e.j >> 5], 1u << (e.j & 31)); else atomicAnd(&mask_s[e.j >> 5], ~(1u << (e.j & 31)));
        }
        if (b > n_fin) { n_changes += b - n_fin; n_fin = b; }
        __syncwarp();
    };
    for (;;) {
        if (in_flight && st_done(ctl, seq)) {
            STP(1);
            const int f = ctl.r_fail;
            in_flight = false; idle_resolved = 0;
            if (f == ST_NONE) {
                n_com += fl_n; buf ^= 1; V = fl_hi;       // (its changes are made final below, off the streamers' critical path)
            } else {
                // changes of the pass before f are final; everything decided beyond them is void
                const bool before = lane < fl_n && ctl.ring[(n_com + lane) & (ST_RING - 1)].pos < f;
                const int m = __popc(__ballot_sync(FULL, before));
                finalize(n_com + m);
                n_dec = n_com + m;
                const int slot = n_dec > 0 ? ((n_dec - 1) & (ST_RING - 1)) : ST_RING;
                qa0 = snap[slot * ST_NA + lane]; qa1 = snap[slot * ST_NA + lane + 32];
                rr = n_dec > 0 ? ctl.ring[(n_dec - 1) & (ST_RING - 1)].rr : ctl.rr_init;
                ia = __popc(__ballot_sync(FULL, mp0 < f)) + __popc(__ballot_sync(FULL, mp1 < f));
                // the change at f, computed by its owner among the streamers
                const int fj = ctl.f_j;
                const double delta = ctl.f_delta, t = ctl.f_t, xl2 = ctl.f_xl2, bnew = ctl.f_bnew;
                const double g0 = mj0 >= 0 ? gram_entry<DESIGN>(P, fj, mj0) : 0.0;
                const double g1 = mj1 >= 0 ? gram_entry<DESIGN>(P, fj, mj1) : 0.0;
                rr = rr - 2.0 * delta * t + delta * delta * xl2;
                qa0 = fma(-delta, g0, qa0); qa1 = fma(-delta, g1, qa1);
                const int e = n_dec & (ST_RING - 1);
                if (lane == 0) {
                    ctl.ring[e].pos = f; ctl.ring[e].j = fj; ctl.ring[e].delta = delta; ctl.ring[e].bnew = bnew; ctl.ring[e].rr = rr;
                    if (DESIGN == BLIPGPU_DESIGN_DENSE)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.G + (int64_t)fj * P.ldg),
                                     "r"((unsigned)(p2 * sizeof(double))) : "memory");
                }
                snap[e * ST_NA + lane] = qa0; snap[e * ST_NA + lane + 32] = qa1;
                __syncwarp();
                ++n_dec;
                finalize(n_dec);
                V = f;
                STP_CNT(10, 1);
            }
            STP(2);
            continue;
        }
        const int ahead = n_dec - n_com;
        const bool can_resolve = ia < T && ahead < ST_LOOK;
        // the streamers are idle: a full list, nothing more to decide, or a few redraws decided while they wait
        if (!in_flight && (ahead >= ST_MAXC || !can_resolve || (ahead > 0 && idle_resolved >= ST_IDLE))) {
            // ---- plan the next pass: the first decided changes, checked up to the next position somebody else decides
            const int nL = min(ST_MAXC, ahead);
            if (nL == 0 && V >= p - 1) break;
            int hi;
            if (n_com + nL < n_dec) hi = ctl.ring[(n_com + nL) & (ST_RING - 1)].pos - 1;
            else if (ia < T) hi = ctl.srt_pos[ia] - 1;
            else hi = p - 1;
            if (hi < V) hi = V;
            if (lane < nL) {
                const StChange& e = ctl.ring[(n_com + lane) & (ST_RING - 1)];
                ps->j[lane] = e.j; ps->pos[lane] = e.pos; ps->d[lane] = e.delta;
            }
            if (lane == 0) { ps->n = nL; ps->lo = V; ps->hi = hi; ps->src = buf; ps->s_fail = ST_NONE; }
            st_issue(ps, ST_CMD_PASS, seq);
            in_flight = true; fl_n = nL; fl_hi = hi; ++n_passes;
            idle_resolved = 0;
            STP_CNT(11, nL);
            STP(3);
            continue;
        }
        if (n_fin < n_com) { finalize(n_com); STP(2); continue; }
        if (can_resolve) {
            // ---- the next tracked redraw, assuming nothing else changes before it
            const int a = ia;
            const int pos = ctl.srt_pos[a], j = ctl.srt_j[a];
            const double q = __shfl_sync(FULL, (a & 32) ? qa1 : qa0, a & 31);
            const double bold = ctl.srt_bold[a], xl2 = ctl.srt_xl2[a];
            const double XjTr = q + bold * xl2;
            const bool spike = wb_spike(XjTr, xl2, ctl.srt_fA[a], ctl.srt_fc1[a], ctl.srt_fL[a], h, u_s, key, pos, sweep);
            double bnew = 0.0;
            if (!spike) bnew = ctl.srt_zpx[a] + ctl.srt_zpy[a] * XjTr / h.sigma2;
            const double delta = bnew - bold;
            rr = rr - 2.0 * delta * q + delta * delta * xl2;
            if (mj0 >= 0) qa0 = fma(-delta, GA[a * ST_NA + lane], qa0);
            if (mj1 >= 0) qa1 = fma(-delta, GA[a * ST_NA + lane + 32], qa1);
            const int e = n_dec & (ST_RING - 1);
            if (lane == 0) {
                ctl.ring[e].pos = pos; ctl.ring[e].j = j; ctl.ring[e].delta = delta; ctl.ring[e].bnew = bnew; ctl.ring[e].rr = rr;
            }
            snap[e * ST_NA + lane] = qa0; snap[e * ST_NA + lane + 32] = qa1;
            __syncwarp();
            ++n_dec; ++ia; ++idle_resolved;
            STP_CNT(12, 1);
            STP(4);
            continue;
        }
        __nanosleep(20);
    }
    // ---- sweep done: every change is committed in q[buf] and final
    finalize(n_dec);
    if (lane == 0) { ctl.buf = buf; ctl.sh.rr = rr; }
    __syncwarp();
    STP(1);
}

template <int DESIGN>
__global__ void __launch_bounds__(ST_THREADS, 1) gibbs_gram_stream_kernel(SweepParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double sm[];
    __shared__ StCtl ctl;
    __shared__ StPass pass;

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();                // 0 = control, 1 = stream
    const int tid = threadIdx.x;
    const int c = blockIdx.x >> 1;
    const int p = P.p, p2 = P.p2;
    double* qbuf = sm;                                         // [2][p2]            (stream CTA)
    double* GA = sm + 2 * (size_t)p2;                          // [ST_NA][ST_NA]     (control CTA)
    double* snap = GA + ST_NA * ST_NA;                         // [ST_RING + 1][ST_NA]
    uint32_t* mask_s = (uint32_t*)(snap + (ST_RING + 1) * ST_NA);   // control CTA: the active set; stream CTA: its copy
    uint32_t* trk_s = mask_s + P.words;
    StCtl* rc = cluster.map_shared_rank(&ctl, 0);
    StPass* ps = cluster.map_shared_rank(&pass, 1);
    const double* q_r = cluster.map_shared_rank(qbuf, 1);

    if (rank == 1) {
        const double* q_g = P.q + (int64_t)c * p2;
        for (int k = tid; k < p2; k += ST_THREADS) qbuf[k] = __ldcg(q_g + k);
        if (tid == 0) { pass.cmd_seq = 0; pass.cmd_type = 0; pass.s_cmd = 0; pass.s_fail = ST_NONE; pass.src = 0; }
    } else {
        for (int w = tid; w < P.words; w += ST_THREADS) { mask_s[w] = __ldcg(P.mask + (int64_t)c * P.words + w); trk_s[w] = 0u; }
        if (tid == 0) {
            ctl.sh.sigma2 = __ldcg(P.hstate + GB_HS * c + 0); ctl.sh.tau2 = __ldcg(P.hstate + GB_HS * c + 1);
            ctl.sh.p0 = __ldcg(P.hstate + GB_HS * c + 2); ctl.sh.rr = __ldcg(P.hstate + GB_HS * c + 3);
            ctl.pre_ready = 0; ctl.d_sweep = 0; ctl.done_seq = 0; ctl.r_fail = ST_NONE; ctl.buf = 0;
        }
    }
    cluster.sync();

    if (rank == 1) {
        st_streamer_cta<DESIGN>(P, c, pass, rc, qbuf, mask_s, trk_s, cluster.map_shared_rank(mask_s, 0), cluster.map_shared_rank(trk_s, 0));
    } else if (tid < ST_NP) {
        st_precompute_role(P, c, ctl);
    } else {
        const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
        double* beta_g = P.beta + (int64_t)c * p;
        int seq = 0;
        long long n_changes = 0, n_passes = 0;
        STP_DECL
        for (int s = 0; s < P.S; ++s) {
            STP(13);
            st_sweep_start<DESIGN>(P, c, s, ctl, ps, q_r, GA, mask_s, trk_s, seq);
            if (tid >= ST_R0) st_resolver_role<DESIGN>(P, c, s, ctl, ps, q_r, GA, snap, mask_s, seq, n_changes, n_passes STP_PASS);
            SyncE::sync();
            sweep_epilogue<SyncE>(P, &ctl.sh, mask_s, beta_g, ctl.scratch, key, c, s, ctl.sh.rr, nullptr, false);
            SyncE::sync();
        }
        STP(13);
        // ---- launch end: the chain state back to global memory
        if (tid >= ST_R0) {
            if ((tid & 31) == 0) ps->src = ctl.buf;
            st_issue(ps, ST_CMD_QUIT, seq);
        }
        for (int w = SyncE::tid(); w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
        if (tid == ST_R0) {
            P.hstate[GB_HS * c + 0] = ctl.sh.sigma2; P.hstate[GB_HS * c + 1] = ctl.sh.tau2;
            P.hstate[GB_HS * c + 2] = ctl.sh.p0; P.hstate[GB_HS * c + 3] = ctl.sh.rr;
            P.hstate[GB_HS * c + 4] = __ldcg(P.hstate + GB_HS * c + 4) + (double)n_changes;
            P.hstate[GB_HS * c + 5] = __ldcg(P.hstate + GB_HS * c + 5) + (double)n_passes;
        }
#ifdef BLIP_PHASE_CLOCKS
        if (c == 0 && tid == ST_R0)
            printf("st resolver | sweepstart %lld wait-pass %lld handle+finalize %lld issue %lld resolve %lld epilogue %lld | fails %lld columns %lld resolved %lld passes %lld changes %lld\n",
                   sp_t[0], sp_t[1], sp_t[2], sp_t[3], sp_t[4], sp_t[13], sp_t[10], sp_t[11], sp_t[12], n_passes, n_changes);
#endif
    }
    // a CTA's shared memory must outlive the peer's last access to it
    cluster.sync();
}
