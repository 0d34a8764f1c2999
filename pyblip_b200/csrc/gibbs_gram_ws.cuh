// gibbs_gram_ws.cuh -- gram-form sweep for FEW chains: role-specialised, one or two SMs per chain.
//
// With no more chains than SMs (north_star's 512 chains sharded over 4 or 8 GPUs) the sweep of
// gibbs_gram_kernel is one latency chain per chain: a third of it is the streaming of Gram
// columns into q (the flush), a sixth the state-independent per-sweep precompute (visiting order,
// logits of the spike uniforms), and neither depends on the decisions being taken.  Here the two
// kinds of work run concurrently:
//
//   D  decision warps (8): the sliding 256-position speculation window of gibbs_gram_kernel,
//      unchanged in arithmetic (same guarded fast decision, same per-thread q_pos register
//      corrected by one gathered Gram entry per change); they never touch a Gram column.
//   F  helper warps (8): (1) apply the changes D publishes on a ring to q -- q lives in shared
//      memory TWICE; F reads buffer e, subtracts the next batch of columns in change order and
//      writes buffer e^1, then publishes (epoch, number of changes applied); (2) in between,
//      precompute the next sweep's visiting order and logits into the other half of a
//      double-buffered scratch.
//
// Two placements of the roles (template parameter CLUSTER):
//   false  one 512-thread CTA per chain: D = warps 0-7, F = warps 8-15 of the same SM
//          (chains <= number of SMs);
//   true   a 2-CTA thread-block cluster per chain: CTA 0 is D, CTA 1 is F and owns q; D reads q
//          and writes the ring through distributed shared memory.  The column stream then runs on
//          an SM of its own: on one SM the helpers' bursts of 16-byte loads sit in the same L1
//          pipeline as the decision warps' gathers and slow them down by about as much as the
//          flush used to cost (measured) -- (chains <= number of SMs / 2).
//
// A position that enters a D thread's ownership reads q from the latest published buffer and
// replays the ring entries that buffer does not contain yet (at most WS_RMAX gathered entries).
// D adopts a new epoch only at a barrier and acknowledges it afterwards, and F does not overwrite
// a buffer before the epoch after it has been acknowledged, so a buffer is immutable while any D
// thread may read it.  Every element of q sees the same fma sequence, in change order, as in
// gibbs_gram_kernel: results are bit-identical to it (tests/test_ws_gpu.py).
//
// Reference semantics: /root/reference/pyblip/linear/_linear.pyx:126-229 (unchanged).
#pragma once
#include <cooperative_groups.h>
#include "gibbs_kernels.cuh"

constexpr int WS_THREADS = 2 * GB_BLOCK;   // single-SM placement: 8 decision warps + 8 helper warps
constexpr int WS_RING = 16;                // ring slots (power of two)
#ifndef BLIP_WS_RMAX
#define BLIP_WS_RMAX 4
#endif
constexpr int WS_RMAX = BLIP_WS_RMAX;      // most changes ahead of the published buffer
constexpr int WS_PEND = 2;                 // changes applied per pass over q
#ifndef BLIP_WS_SLOTS
#define BLIP_WS_SLOTS 6
#endif
constexpr int WS_SLOTS = BLIP_WS_SLOTS;    // depth of a helper thread's rotating load pipeline (16-byte loads per column)
constexpr int WS_PRE_CHUNK = 2 * GB_BLOCK; // positions precomputed per helper action

#ifdef BLIP_PHASE_CLOCKS
#define PHF_DECL long long ph_t[12] = {0,0,0,0,0,0,0,0,0,0,0,0}; long long ph_last = clock64();
#define PHF(i) do { if (ftid == 0) { long long _n = clock64(); ph_t[i] += _n - ph_last; ph_last = _n; } } while (0)
#else
#define PHF_DECL
#define PHF(i) do {} while (0)
#endif
typedef SyncNamed<1, GB_BLOCK> SyncD;
typedef SyncNamed<2, GB_BLOCK> SyncF;

// Control words.  Every field has ONE reader role, which polls its own CTA's copy; the writer
// stores into the reader's copy (the same struct in the single-SM placement, the peer CTA's through
// distributed shared memory in the cluster placement).
struct WsCtl {
    unsigned long long pub;        // F -> D   (epoch << 32) | changes applied in buffer (epoch & 1)
    unsigned long long slot[3];    // D internal: pub as adopted at a barrier (rotating)
    int ring_t;                    // D -> F   changes appended so far
    int ack_epoch;                 // D -> F   epoch every D thread has adopted
    int d_sweep;                   // D -> F   sweep (launch-local) D is running
    int pre_ready;                 // F -> D   sweeps < pre_ready have their scratch
    int quit;                      // D -> F
    int f_act, f_n;                // F internal
    int ring_j[WS_RING];           // D -> F   (D keeps its own copy for the replays)
    double ring_d[WS_RING];
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
__device__ __forceinline__ int ld_volatile_int(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ void st_volatile_int(int* p, int v) { *(volatile int*)p = v; }
// orders this thread's earlier accesses before its later ones, at the scope the two roles share
template <bool CLUSTER> __device__ __forceinline__ void ws_fence()
{
    if (CLUSTER) asm volatile("fence.acq_rel.cluster;" ::: "memory");
    else __threadfence_block();
}

// dst = src - sum_k pd[k] * G[:, pj[k]] (k < n, in order), helper warps only (ftid in [0, 256)).
// (A variant that streamed the columns with cp.async.bulk into shared-memory stages was measured at
// ~500 cycles per bulk copy whatever its size, 3x slower than these loads; left out.)
template <int DESIGN>
__device__ __forceinline__ void ws_flush(const SweepParams& P, const double* __restrict__ src, double* __restrict__ dst,
                                         const int (&pj)[WS_PEND], const double (&pd)[WS_PEND], int n, int ftid,
                                         double* dst2 = nullptr)
{
    // A rotating register pipeline of WS_SLOTS groups, one 16-byte load per column and group: every
    // step applies the oldest group and re-issues its loads WS_SLOTS steps ahead, so the helpers feed
    // the L1 pipeline two loads at a time instead of in bursts of twenty -- the decision warps' gathers
    // share that pipeline.
    const int p2 = P.p2;
    constexpr int STEP = 2 * GB_BLOCK;                  // doubles covered by one group step
    double2 g[WS_SLOTS][WS_PEND];
    auto load = [&](double2 (&gg)[WS_PEND], int kk) {
#pragma unroll
        for (int k = 0; k < WS_PEND; ++k)
            gg[k] = (kk < p2 && k < n) ? gram_pair<DESIGN>(P, pj[k], kk) : make_double2(0.0, 0.0);
    };
    auto apply = [&](const double2 (&gg)[WS_PEND], int kk) {
        if (kk < p2) {
            double2 qq = *reinterpret_cast<const double2*>(src + kk);
#pragma unroll
            for (int k = 0; k < WS_PEND; ++k)
                if (k < n) { qq.x = fma(-pd[k], gg[k].x, qq.x); qq.y = fma(-pd[k], gg[k].y, qq.y); }
            *reinterpret_cast<double2*>(dst + kk) = qq;
            if (dst2) *reinterpret_cast<double2*>(dst2 + kk) = qq;
        }
    };
    int k0 = 2 * ftid;
#pragma unroll
    for (int i = 0; i < WS_SLOTS; ++i) load(g[i], k0 + i * STEP);
    for (; k0 < p2; k0 += WS_SLOTS * STEP) {
#pragma unroll
        for (int i = 0; i < WS_SLOTS; ++i) {
            apply(g[i], k0 + i * STEP);
            load(g[i], k0 + (i + WS_SLOTS) * STEP);
        }
    }
}

// =================================================================================================
// helper role: `mine` is polled, `peer` is D's copy of the control words; qbuf is local
// =================================================================================================
template <int DESIGN, bool CLUSTER>
__device__ __forceinline__ void ws_helper_role(const SweepParams& P, int c, int ftid, double* qbuf, WsCtl* mine, WsCtl* peer,
                                               double* q_peer = nullptr)
{
    const int p = P.p, p2 = P.p2, s_end = P.S;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    const bool keyed_perm = P.perm == nullptr;
    int applied = 0, epoch = 0, pre_next = 0, pre_pos = 0;
    PermKey pk;
    PHF_DECL
    for (;;) {
        PHF(0);
        if (ftid == 0) {
            int act, n = 0;
            for (;;) {
                const int t = ld_volatile_int(&mine->ring_t);
                if (t > applied && ld_volatile_int(&mine->ack_epoch) == epoch) { act = 1; n = min(t - applied, WS_PEND); break; }
                if (pre_next < s_end && pre_next <= ld_volatile_int(&mine->d_sweep) + 1) { act = 2; break; }
                if (t == applied && ld_volatile_int(&mine->quit)) { act = 3; break; }
            }
            ws_fence<CLUSTER>();
            mine->f_act = act; mine->f_n = n;
        }
        SyncF::sync();
        const int act = mine->f_act, n = mine->f_n;
        PHF(1);
        if (act == 3) break;
        if (act == 1) {
            int pj[WS_PEND]; double pd[WS_PEND];
#pragma unroll
            for (int k = 0; k < WS_PEND; ++k) {
                const int slot = (applied + k) & (WS_RING - 1);
                pj[k] = k < n ? mine->ring_j[slot] : 0;
                pd[k] = k < n ? mine->ring_d[slot] : 0.0;
            }
            // (q_peer: a second copy of q in the decision CTA's shared memory, gibbs_gram_ws2.cuh -- the new buffer is
            //  pushed into it with distributed-shared-memory stores, ordered before `pub` by the fence below)
            ws_flush<DESIGN>(P, qbuf + (size_t)(epoch & 1) * p2, qbuf + (size_t)((epoch + 1) & 1) * p2, pj, pd, n, ftid,
                             q_peer ? q_peer + (size_t)((epoch + 1) & 1) * p2 : nullptr);
            applied += n;
            ++epoch;
            SyncF::sync();
            if (ftid == 0) {
                ws_fence<CLUSTER>();
                *(volatile unsigned long long*)&peer->pub = ((unsigned long long)(unsigned)epoch << 32) | (unsigned)applied;
            }
            PHF(2);
        } else {
            // state-independent quantities of sweep pre_next, into half (pre_next & 1) of the scratch
            const int sweep = P.sweep0 + pre_next;
            const int64_t half = (int64_t)(pre_next & 1) * P.scr_stride + (int64_t)c * p;
            if (pre_pos == 0 && keyed_perm) pk = perm_key(P.k0, P.k1, key.chain, (uint32_t)sweep, (uint32_t)p);
            const double* u_s = P.u ? P.u + ((int64_t)c * P.S + pre_next) * p : nullptr;
#pragma unroll
            for (int i = 0; i < WS_PRE_CHUNK / GB_BLOCK; ++i) {
                const int pos = pre_pos + ftid + i * GB_BLOCK;
                if (pos < p) {
                    if (keyed_perm) {
                        const int j = (int)perm_at(pk, (uint32_t)pos, (uint32_t)p);
                        if (P.perm16) reinterpret_cast<uint16_t*>(P.scrP)[half + pos] = (uint16_t)j;
                        else reinterpret_cast<int32_t*>(P.scrP)[half + pos] = j;
                    }
                    const double u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                    const bool ex = !(u > 1e-4 && u < 1.0 - 1e-4);
                    P.scrL[half + pos] = ex ? __int_as_float(0x7fc00000) : __logf((float)u) - __logf((float)(1.0 - u));
                }
            }
            pre_pos += WS_PRE_CHUNK;
            if (pre_pos >= p) {
                __threadfence();                              // the scratch is read through L2 by the D warps
                SyncF::sync();
                ++pre_next;
                pre_pos = 0;
                if (ftid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->pre_ready, pre_next); }
            }
            PHF(3);
        }
        SyncF::sync();                                        // f_act / f_n may be rewritten
    }
    // launch end: D has drained the ring; the published buffer goes back to HBM
    {
        const double* q = qbuf + (size_t)(epoch & 1) * p2;
        double* q_g = P.q + (int64_t)c * p2;
        for (int k = ftid; k < p2; k += GB_BLOCK) q_g[k] = q[k];
    }
#ifdef BLIP_PHASE_CLOCKS
    if (ftid == 0 && c == 0)
        printf("ws helper | loop %lld poll %lld flush %lld precompute %lld\n", ph_t[0], ph_t[1], ph_t[2], ph_t[3]);
#endif
}

// shared memory of the decision role
struct WsDecideShared {
    SweepShared sh;
    GramBcast bc;
    int s_first[3];
    int sj[GB_BLOCK];
    int act_pos[GB_BLOCK];
    int act_n;
    double scratch[40];
};

// =================================================================================================
// decision role: `mine` is polled, `peer` is F's copy of the control words; q_d are the two q
// buffers as seen from here (the peer CTA's shared memory in the cluster placement)
// =================================================================================================
template <int DESIGN, bool CLUSTER>
__device__ __forceinline__ void ws_decision_role(const SweepParams& P, int c, double* q_d, uint32_t* mask_s, WsCtl* mine,
                                                 WsCtl* peer, WsDecideShared& ds)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = P.p, p2 = P.p2, s_end = P.S;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    SweepShared& sh = ds.sh;
    int* s_first = ds.s_first;
    int* sj = ds.sj;
    double* beta_g = P.beta + (int64_t)c * p;
    int parity = 0, eparity = 0;
    int epoch = 0, applied = 0, t_chg = 0;           // block-uniform
    // adopt the helpers' latest (epoch, applied) -- one barrier; thread 0 acknowledges afterwards
    auto adopt = [&]() {
        if (tid == 0) { const unsigned long long v = ld_volatile_u64(&mine->pub); ws_fence<CLUSTER>(); mine->slot[eparity] = v; }
        SyncD::sync();
        const unsigned long long v = mine->slot[eparity];
        eparity = eparity == 2 ? 0 : eparity + 1;
        epoch = (int)(v >> 32); applied = (int)(unsigned)v;
        if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ack_epoch, epoch); }
    };
    auto drain = [&]() { while (applied != t_chg) adopt(); };   // every change is in the published buffer

    long long n_changes = 0, n_windows = 0;
    PH_DECL
    for (int s = 0; s < s_end; ++s) {
        PH(0);
        const int sweep = P.sweep0 + s;
        const int64_t soff = ((int64_t)c * P.S + s) * p;
        const int64_t half = (int64_t)(s & 1) * P.scr_stride + (int64_t)c * p;
        const PermView perm_s = P.perm ? perm_view(P, c, s) : PermView{ P.scrP, half, P.perm16 };
        const double* u_s = P.u ? P.u + soff : nullptr;
        const double* z_s = P.z ? P.z + soff : nullptr;
        if (tid == 0) {
            st_volatile_int(&peer->d_sweep, s);
            while (ld_volatile_int(&mine->pre_ready) <= s) { }
            ws_fence<CLUSTER>();
            __threadfence();                                  // the helpers wrote the scratch to global memory
        }
        PH(1);
        // ---- periodic rebuild of q and ||r||^2 from q0 = X^T y (bounds rounding drift)
        if (P.gram_refresh > 0 && sweep > 0 && sweep % P.gram_refresh == 0) {
            drain();
            double* q = q_d + (size_t)(epoch & 1) * p2;      // the helpers are idle: nothing is pending
            for (int k = tid; k < p2; k += GB_BLOCK) q[k] = k < p ? P.q0[k] : 0.0;
            SyncD::sync();
            for (int w = 0; w < P.words; ++w) {
                uint32_t m = mask_s[w];
                while (m) {
                    int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    gram_axpy<DESIGN>(P, q, j, __ldcg(beta_g + j));
                }
            }
            SyncD::sync();
            double acc = 0.0;
            for (int w = tid; w < P.words; w += GB_BLOCK) {
                uint32_t m = mask_s[w];
                while (m) {
                    int j = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    acc = fma(__ldcg(beta_g + j), P.q0[j] + q[j], acc);
                }
            }
            acc = block_sum<SyncD>(acc, ds.scratch);
            if (tid == 0) sh.rr = P.yy - acc;
            ws_fence<CLUSTER>();                              // the rebuilt q before the next ring_t the helpers see
        }
        if (tid == 0) { sh.logodds = log(sh.p0) - log(1 - sh.p0); ds.act_n = 0; }
        SyncD::sync();
        const Hyper h = { sh.sigma2, sh.tau2, sh.logodds };
        const double ts = h.tau2 / h.sigma2;
        const float* Ls = P.scrL + half;
        double2* Zs = P.scrZ + (int64_t)c * p;

        // ---- the state-dependent part of the sweep start: the slab draws of the active coordinates
        for (int base = tid; base < p; base += 4 * GB_BLOCK) {
            int jj[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const int pos = base + i * GB_BLOCK; jj[i] = pos < p ? perm_s[pos] : -1; }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = jj[i];
                if (j >= 0 && ((mask_s[j >> 5] >> (j & 31)) & 1u)) {
                    const int slot = atomicAdd(&ds.act_n, 1);
                    if (slot < GB_BLOCK) ds.act_pos[slot] = base + i * GB_BLOCK;
                    else slab_prepare(P, Zs, perm_s, h, key, z_s, base + i * GB_BLOCK, sweep);
                }
            }
        }
        SyncD::sync();
        if (tid < min(ds.act_n, GB_BLOCK)) slab_prepare(P, Zs, perm_s, h, key, z_s, ds.act_pos[tid], sweep);
        adopt();                                                 // (also the barrier after slab_prepare)
        PH(2);

        int c_pos = -1, c_j = 0, n_j = 0, nn_j = 0;
        bool c_act = false;
        int n_def = 0;
        double d_g[WS_RMAX], d_d[WS_RMAX];
        float f_A = 0.f, f_c1 = 0.f, f_L = 0.f, n_L = 0.f;
        double c_xl2 = 0.0, n_xl2 = 0.0, c_bold = 0.0;
        double2 c_zp = make_double2(0.0, 0.0);
        double q_pos = 0.0;
        const float f_logodds = (float)h.logodds, f_tau2 = (float)h.tau2, f_sigma2 = (float)h.sigma2;
        const float f_ts = (float)ts;
        auto fill = [&](int pos, int jnow, bool publish = true) -> double {
            c_pos = pos;
            c_j = n_j;
            if (publish) sj[tid] = n_j;
            q_pos = q_d[(size_t)(epoch & 1) * p2 + c_j];        // (a distributed-shared-memory load in the cluster placement)
            const double gnow = jnow >= 0 ? gram_entry<DESIGN>(P, jnow, c_j) : 0.0;
            const int npend = t_chg - applied;                   // <= WS_RMAX
#pragma unroll
            for (int k = 0; k < WS_RMAX; ++k)
                if (k < npend) {
                    const int slot = (applied + k) & (WS_RING - 1);
                    d_g[k] = gram_entry<DESIGN>(P, mine->ring_j[slot], c_j); d_d[k] = mine->ring_d[slot];
                }
            n_def = npend;
            c_xl2 = n_xl2;
            f_L = n_L;
            c_act = (mask_s[c_j >> 5] >> (c_j & 31)) & 1u;
            if (c_act) { c_bold = __ldcg(beta_g + c_j); c_zp = __ldcg(Zs + pos); } else c_bold = 0.0;
            n_j = nn_j;
            if (pos + GB_BLOCK < p) { n_xl2 = P.Xl2[n_j]; n_L = __ldcg(Ls + pos + GB_BLOCK); }
            if (pos + 2 * GB_BLOCK < p) nn_j = perm_s[pos + 2 * GB_BLOCK];
            const float xl2f = (float)c_xl2;
            f_A = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
            f_c1 = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
            return gnow;
        };
        auto replay = [&]() {
#pragma unroll
            for (int k = 0; k < WS_RMAX; ++k) if (k < n_def) q_pos = fma(-d_d[k], d_g[k], q_pos);
            n_def = 0;
        };
        if (tid < p) {
            n_j = perm_s[tid]; n_xl2 = P.Xl2[n_j]; n_L = __ldcg(Ls + tid);
            if (tid + GB_BLOCK < p) nn_j = perm_s[tid + GB_BLOCK];
            fill(tid, -1);
        }
        PH(3);

        int pos0 = 0;
        while (pos0 < p) {
            ++n_windows;
            const int off = (tid - pos0) & (GB_BLOCK - 1);
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
                XjTr = c_act ? q_pos + c_bold * c_xl2 : q_pos;
                const double x2c = (double)f_c1 * (XjTr * XjTr);
                const double E = (double)f_A - x2c;
                const double margin = E - (double)f_L;
                const double guard = 1e-4 * (1.0 + fabs((double)f_A) + x2c + fabs((double)f_L));
                if (!(fabs(margin) > guard) || !(fmax(E, (double)f_L) < 20.0)) {
                    const double c_u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
                    spike = (c_u <= kappa_of(XjTr, c_xl2, h));
                } else {
                    spike = margin > 0.0;
                }
                change = c_act || !spike;
            }
            const unsigned m = __ballot_sync(0xffffffffu, change);
            const int r = pos0 & (GB_BLOCK - 1);
            if (m && lane == 0) {
                int l;
                if (warp == (r >> 5)) {
                    const unsigned hi = m & (0xffffffffu << (r & 31));
                    l = hi ? __ffs(hi) - 1 : __ffs(m) - 1;
                } else l = __ffs(m) - 1;
                atomicMin(&s_first[parity], (32 * warp + l - r) & (GB_BLOCK - 1));
            }
            PH(4);
            adopt();                                          // the window's barrier; newest buffer from here on
            PH(5);
            int first = s_first[parity];
            if (first >= GB_BLOCK) first = -1;
            if (tid == 0) s_first[parity == 0 ? 2 : parity - 1] = GB_BLOCK;
            parity = parity == 2 ? 0 : parity + 1;
            if (first < 0) {
                pos0 += GB_BLOCK;
                if (pos + GB_BLOCK < p) fill(pos + GB_BLOCK, -1); else c_pos = -1;
                PH(6);
                continue;
            }
            // a change is about to go on the ring: wait if the helpers are WS_RMAX changes behind
            while (t_chg - applied >= WS_RMAX) adopt();
            PH(7);
            ++n_changes;
            const int jstar = sj[(pos0 + first) & (GB_BLOCK - 1)];
            double gj = 0.0;
            if (off > first) {
                if (c_pos >= 0) gj = gram_entry<DESIGN>(P, jstar, c_j);
            } else if (off == first) {
                double bnew = 0.0;
                if (!spike) {
                    if (c_act) bnew = c_zp.x + c_zp.y * XjTr / h.sigma2;
                    else bnew = slab_draw(XjTr, c_xl2, h, z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0));
                }
                const double delta = bnew - c_bold;
                ds.bc.delta = delta;
                const int slot = t_chg & (WS_RING - 1);
                mine->ring_j[slot] = c_j; mine->ring_d[slot] = delta;
                if (CLUSTER) { peer->ring_j[slot] = c_j; peer->ring_d[slot] = delta; }
                sh.rr = sh.rr - 2.0 * delta * q_pos + delta * delta * c_xl2;
                __stcg(beta_g + c_j, bnew);
                if (bnew != 0) mask_s[c_j >> 5] |= (1u << (c_j & 31));
                else mask_s[c_j >> 5] &= ~(1u << (c_j & 31));
                // a column that is not L2-resident is on its way by the time the helpers stream it
                if (DESIGN == BLIPGPU_DESIGN_DENSE)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.G + (int64_t)c_j * P.ldg),
                                 "r"((unsigned)(p2 * sizeof(double))) : "memory");
                ws_fence<CLUSTER>();
            }
            if (off <= first) {
                if (pos + GB_BLOCK < p) gj = fill(pos + GB_BLOCK, jstar, off != first); else c_pos = -1;
            }
            PH(8);
            SyncD::sync();
            PH(9);
            if (off == first && c_pos >= 0) sj[tid] = c_j;
            replay();
            q_pos = fma(-ds.bc.delta, gj, q_pos);
            ++t_chg;
            // (the ring entry was written, and fenced, by its owner before the barrier)
            if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ring_t, t_chg); }
            pos0 += first + 1;
            PH(10);
        }
        SyncD::sync();
        sweep_epilogue<SyncD>(P, &sh, mask_s, beta_g, ds.scratch, key, c, s, sh.rr, nullptr, false);
        SyncD::sync();
    }
    PH(0);
    // ---- launch end: everything into the published buffer (the helpers write it back to HBM)
    drain();
    if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->quit, 1); }
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2;
        P.hstate[GB_HS * c + 2] = sh.p0; P.hstate[GB_HS * c + 3] = sh.rr;
        P.hstate[GB_HS * c + 4] = __ldcg(P.hstate + GB_HS * c + 4) + (double)n_changes;
        P.hstate[GB_HS * c + 5] = __ldcg(P.hstate + GB_HS * c + 5) + (double)n_windows;
    }
#ifdef BLIP_PHASE_CLOCKS
    if (tid == 0 && c == 0)
        printf("ws decide | epilogue %lld wait-pre %lld sweepstart %lld fill0 %lld | eval %lld sync1 %lld nullfill %lld throttle %lld "
               "owner/fill/gather %lld sync2 %lld apply %lld\n", ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5],
               ph_t[6], ph_t[7], ph_t[8], ph_t[9], ph_t[10]);
#endif
}

template <int DESIGN, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? GB_BLOCK : WS_THREADS, 1) gibbs_gram_ws_kernel(SweepParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double sm[];
    __shared__ WsCtl ctl;
    __shared__ WsDecideShared ds;

    const int p2 = P.p2;
    double* qbuf = sm;                                // [2][p2]  (cluster placement: used in the helper CTA only)
    uint32_t* mask_s = (uint32_t*)(sm + 2 * (size_t)p2);
    int c, role, rtid;                                // role 0 = decision, 1 = helper
    WsCtl* peer = &ctl;
    double* q_d = qbuf;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        role = (int)cluster.block_rank();
        c = blockIdx.x >> 1;
        rtid = threadIdx.x;
        peer = cluster.map_shared_rank(&ctl, role ^ 1);
        q_d = cluster.map_shared_rank(qbuf, 1);
    } else {
        role = threadIdx.x >= GB_BLOCK;
        c = blockIdx.x;
        rtid = threadIdx.x & (GB_BLOCK - 1);
    }

    // ---- chain state -> shared memory: q where the helpers are, the rest where the decisions are
    const int nload = CLUSTER ? GB_BLOCK : WS_THREADS;
    if (!CLUSTER || role == 1) {
        const double* q_g = P.q + (int64_t)c * p2;
        for (int k = threadIdx.x; k < p2; k += nload) qbuf[k] = __ldcg(q_g + k);
    }
    if (!CLUSTER || role == 0) {
        for (int w = threadIdx.x; w < P.words; w += nload) mask_s[w] = __ldcg(P.mask + (int64_t)c * P.words + w);
    }
    if (threadIdx.x == 0) {
        ds.sh.sigma2 = __ldcg(P.hstate + GB_HS * c + 0); ds.sh.tau2 = __ldcg(P.hstate + GB_HS * c + 1);
        ds.sh.p0 = __ldcg(P.hstate + GB_HS * c + 2); ds.sh.rr = __ldcg(P.hstate + GB_HS * c + 3);
        ctl.pub = 0ull; ctl.ring_t = 0; ctl.ack_epoch = 0; ctl.d_sweep = 0; ctl.pre_ready = 0; ctl.quit = 0;
        ctl.f_act = 0; ctl.f_n = 0;
        ds.s_first[0] = ds.s_first[1] = ds.s_first[2] = GB_BLOCK;
    }
    if (CLUSTER) cg::this_cluster().sync();           // both CTAs' control words exist before anybody writes into the peer's
    else __syncthreads();

    if (role == 1) ws_helper_role<DESIGN, CLUSTER>(P, c, rtid, qbuf, &ctl, peer);
    else ws_decision_role<DESIGN, CLUSTER>(P, c, q_d, mask_s, &ctl, peer, ds);

    // a CTA's shared memory must outlive the peer's last access to it
    if (CLUSTER) cg::this_cluster().sync();
}
