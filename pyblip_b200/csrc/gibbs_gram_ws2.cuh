// gibbs_gram_ws2.cuh -- the role-specialised gram kernel (gibbs_gram_ws.cuh) with the L2 round trips taken
// off the decision warps' latency chain.
//
// In gibbs_gram_ws_kernel a change costs the decision warps two exposed L2 latencies: the Gram entry
// G[j*][c_j] that corrects every thread's q_pos can only be requested once j* is known, and a position that
// enters a thread's ownership reads q from shared memory and gathers the entries of the pending changes right
// before its first evaluation.  Measured (phase clocks, 64 chains at C2): 1600 cycles per change waiting for
// those gathers, of ~3500 per change in all.  Here
//
//   * every thread tracks TWO positions, the one it owns and the one it will own next (256 further on): both
//     receive every change as it happens (one fma each), so a position is complete -- q, pending changes,
//     coefficient, slab draw, logit -- a whole ownership period before it is first evaluated, and a refill is
//     a register move;
//   * the Gram entries of the NEXT change are requested one step ahead: the positions of the coordinates that
//     are active at sweep start are known, every one of them is a change when it is visited, and ~ 90 % of a
//     stationary sweep's changes are exactly these.  The list is sorted at sweep start; after every change the
//     threads request G[a][c_j] and G[a][n_j] for the next listed coordinate a, and use them when the next
//     change turns out to be a (an activation of a null coordinate in between falls back to the gather-now path).
//
// Both items add scattered loads (a fully divergent warp load holds the L1 tag stage for 32 cycles), and on one SM the
// decision warps share that pipeline with the helpers' column stream: the single-SM placement of this kernel measures
// SLOWER than gibbs_gram_ws_kernel (6.7 vs 5.7 ms per 32 sweeps of 64 chains).  The placement this kernel is meant
// for is the 2-CTA cluster (CLUSTER = true): decisions on one SM, helpers on the other, and -- unlike the cluster
// placement of gibbs_gram_ws_kernel -- q lives on BOTH SMs: the helpers push every new buffer into the decision CTA
// with distributed-shared-memory stores, so the decision warps only ever read local shared memory, and a ninth
// (control) warp of the decision CTA does all cluster-scope fences and remote control-word stores, so that no warp
// with Gram loads in flight ever waits at a fence and the decision warps' L1 is never invalidated.
//
// Arithmetic, decisions, ring / epoch protocol with the helper warps: unchanged; every element of q sees the same
// fma sequence in change order, so results are bit-identical to gibbs_gram_kernel (tests/test_ws_gpu.py).
//
// Reference semantics: /root/reference/pyblip/linear/_linear.pyx:126-229 (unchanged).
#pragma once
#include "gibbs_gram_ws.cuh"

struct Ws2DecideShared {
    SweepShared sh;
    GramBcast bc;
    int s_first[3];
    int sj[GB_BLOCK];
    int act_pos[GB_BLOCK];
    int act_n, na;
    int srt_pos[GB_BLOCK], srt_j[GB_BLOCK];     // the listed active coordinates in visiting order
    double scratch[40];
};

constexpr int WS2_D_THREADS = GB_BLOCK + 32;        // cluster placement: 8 decision warps + the control warp
typedef SyncNamed<3, WS2_D_THREADS> SyncM9;

// `worker` threads (tid < 256) hold positions; in the cluster placement the threads of warp 8 only mirror the
// block-uniform control flow, take part in the main-loop barriers and do the control-word traffic (lane 0).
template <int DESIGN, bool CLUSTER>
__device__ __forceinline__ void ws2_decision_role(const SweepParams& P, int c, double* q_d, uint32_t* mask_s, WsCtl* mine,
                                                  WsCtl* peer, Ws2DecideShared& ds, double* q_f)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = tid < GB_BLOCK;
    const bool ctl_lane = CLUSTER ? tid == GB_BLOCK : tid == 0;
    auto sync_main = [&]() { if (CLUSTER) SyncM9::sync(); else SyncD::sync(); };
    const int p = P.p, p2 = P.p2, s_end = P.S;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    SweepShared& sh = ds.sh;
    int* s_first = ds.s_first;
    int* sj = ds.sj;
    double* beta_g = P.beta + (int64_t)c * p;
    int parity = 0, eparity = 0;
    int epoch = 0, applied = 0, t_chg = 0;           // block-uniform
    // (cluster-scope fences are expensive -- a GPU-scope membar and an L1 invalidate each -- so the control lane
    //  only fences when the helpers have published a new buffer: nothing is acquired or released otherwise)
    auto adopt = [&]() {
        const int old_epoch = epoch;
        if (ctl_lane) {
            const unsigned long long v = ld_volatile_u64(&mine->pub);
            if (!CLUSTER || (int)(v >> 32) != old_epoch) ws_fence<CLUSTER>();
            mine->slot[eparity] = v;
        }
        sync_main();
        const unsigned long long v = mine->slot[eparity];
        eparity = eparity == 2 ? 0 : eparity + 1;
        epoch = (int)(v >> 32); applied = (int)(unsigned)v;
        if (ctl_lane && (!CLUSTER || epoch != old_epoch)) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ack_epoch, epoch); }
    };
    auto drain = [&]() { while (applied != t_chg) adopt(); };

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
        if (ctl_lane) {
            st_volatile_int(&peer->d_sweep, s);
            while (ld_volatile_int(&mine->pre_ready) <= s) { }
            ws_fence<CLUSTER>();
            __threadfence();                                  // the helpers wrote the scratch to global memory
        }
        if (CLUSTER) sync_main();                             // (single SM: thread 0 is a worker and meets the others below)
        PH(1);
        // ---- periodic rebuild of q and ||r||^2 from q0 = X^T y (bounds rounding drift)
        if (P.gram_refresh > 0 && sweep > 0 && sweep % P.gram_refresh == 0) {
            drain();
            if (worker) {
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
                // the helpers' copy takes the rebuilt q too (ordered before the next ring_t by the control lane's fence)
                if (CLUSTER) for (int k = tid; k < p2; k += GB_BLOCK) q_f[(size_t)(epoch & 1) * p2 + k] = q[k];
                else ws_fence<CLUSTER>();
            }
            if (CLUSTER) sync_main();
        }
        if (!worker) {
            // ===== control warp: the block-uniform skeleton of the sweep =====
            adopt();
            int pos0c = 0;
            while (pos0c < p) {
                adopt();
                int first = s_first[parity];
                if (first >= GB_BLOCK) first = -1;
                parity = parity == 2 ? 0 : parity + 1;
                if (first < 0) { pos0c += GB_BLOCK; continue; }
                while (t_chg - applied >= WS_RMAX) adopt();
                sync_main();
                ++t_chg;
                if (ctl_lane) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ring_t, t_chg); }
                pos0c += first + 1;
            }
            continue;
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
        {
            // the listed active coordinates in visiting order (more than the list holds: nothing is predicted)
            const int n_act = ds.act_n;
            if (tid < min(n_act, GB_BLOCK)) {
                const int pos = ds.act_pos[tid];
                slab_prepare(P, Zs, perm_s, h, key, z_s, pos, sweep);
                if (n_act <= GB_BLOCK) {
                    int rank = 0;
                    for (int k = 0; k < n_act; ++k) rank += ds.act_pos[k] < pos;
                    ds.srt_pos[rank] = pos; ds.srt_j[rank] = perm_s[pos];
                }
            }
            if (tid == 0) ds.na = n_act <= GB_BLOCK ? n_act : 0;
        }
        adopt();                                                 // (also the barrier after slab_prepare)
        PH(2);
        const int na = ds.na;

        // ---- per-thread state: the owned position and the one 256 further on
        int c_pos = -1, c_j = 0;
        bool c_act = false;
        float f_A = 0.f, f_c1 = 0.f, f_L = 0.f;
        double c_xl2 = 0.0, c_bold = 0.0, q_pos = 0.0;
        double2 c_zp = make_double2(0.0, 0.0);
        int n_pos = -1, n_j = 0, nn_j = 0, nn_def = 0;
        bool n_act = false;
        float n_L = 0.f;
        double n_xl2 = 0.0, n_bold = 0.0, q_nxt = 0.0;
        double2 n_zp = make_double2(0.0, 0.0);
        double nd_g[WS_RMAX], nd_d[WS_RMAX];
        int pgc_j = -1, pgn_j = -1;                      // coordinates the prefetched Gram entries belong to
        double pg_c = 0.0, pg_n = 0.0;
        const float f_logodds = (float)h.logodds, f_tau2 = (float)h.tau2, f_sigma2 = (float)h.sigma2;
        const float f_ts = (float)ts;
        // everything the position `pos` (coordinate nn_j) needs, from the newest adopted buffer plus the changes
        // that buffer does not contain yet; nothing here is waited for before the next change is applied
        auto init_next = [&](int pos) {
            pgn_j = -1;                                          // a prefetched entry belonged to the previous next position
            if (pos >= p) { n_pos = -1; return; }
            n_pos = pos;
            n_j = nn_j;
            q_nxt = q_d[(size_t)(epoch & 1) * p2 + n_j];
            const int npend = t_chg - applied;                   // <= WS_RMAX
#pragma unroll
            for (int k = 0; k < WS_RMAX; ++k)
                if (k < npend) {
                    const int slot = (applied + k) & (WS_RING - 1);
                    nd_g[k] = gram_entry<DESIGN>(P, mine->ring_j[slot], n_j); nd_d[k] = mine->ring_d[slot];
                }
            nn_def = npend;
            n_xl2 = P.Xl2[n_j];
            n_L = __ldcg(Ls + pos);
            n_act = (mask_s[n_j >> 5] >> (n_j & 31)) & 1u;
            if (n_act) { n_bold = __ldcg(beta_g + n_j); n_zp = __ldcg(Zs + pos); } else n_bold = 0.0;
            if (pos + GB_BLOCK < p) nn_j = perm_s[pos + GB_BLOCK];
        };
        auto replay_next = [&]() {
#pragma unroll
            for (int k = 0; k < WS_RMAX; ++k) if (k < nn_def) q_nxt = fma(-nd_d[k], nd_g[k], q_nxt);
            nn_def = 0;
        };
        // the next position becomes the owned one (`publish`: its coordinate goes on the board right away)
        auto advance = [&](bool publish) {
            replay_next();
            c_pos = n_pos; c_j = n_j;
            if (publish) sj[tid] = c_j;
            q_pos = q_nxt; c_xl2 = n_xl2; f_L = n_L; c_act = n_act; c_bold = n_bold; c_zp = n_zp;
            pg_c = pg_n; pgc_j = pgn_j;
            const float xl2f = (float)c_xl2;
            f_A = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
            f_c1 = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
        };
        int ia = 0, pa = na > 0 ? ds.srt_j[0] : -1;     // block-uniform: next listed change, its coordinate
        if (tid < p) {
            nn_j = perm_s[tid];
            init_next(tid);
            advance(true);
            init_next(tid + GB_BLOCK);
            if (pa >= 0) {
                pg_c = gram_entry<DESIGN>(P, pa, c_j); pgc_j = pa;
                if (n_pos >= 0) { pg_n = gram_entry<DESIGN>(P, pa, n_j); pgn_j = pa; }
            }
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
                    nn_j = perm_s[pos];
                    init_next(pos); advance(true); init_next(pos + GB_BLOCK);
                    pgc_j = -1; pgn_j = -1;
                }
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
                if (pos < p) {
                    if (n_pos >= 0) {
                        advance(true);
                        init_next(c_pos + GB_BLOCK);
                        if (pa >= 0 && n_pos >= 0) { pg_n = gram_entry<DESIGN>(P, pa, n_j); pgn_j = pa; }
                    } else c_pos = -1;
                }
                PH(6);
                continue;
            }
            // a change is about to go on the ring: wait if the helpers are WS_RMAX changes behind
            while (t_chg - applied >= WS_RMAX) adopt();
            PH(7);
            ++n_changes;
            const int jstar = sj[(pos0 + first) & (GB_BLOCK - 1)];
            // the Gram entries of this change for both tracked positions: prefetched if it is the predicted one
            double gc = 0.0, gn = 0.0;
            if (off > first && c_pos >= 0) gc = pgc_j == jstar ? pg_c : gram_entry<DESIGN>(P, jstar, c_j);
            if (n_pos >= 0) gn = pgn_j == jstar ? pg_n : gram_entry<DESIGN>(P, jstar, n_j);
            if (off == first) {
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
                if (DESIGN == BLIPGPU_DESIGN_DENSE)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.G + (int64_t)c_j * P.ldg),
                                 "r"((unsigned)(p2 * sizeof(double))) : "memory");
                if (!CLUSTER) ws_fence<CLUSTER>();
            }
            PH(8);
            sync_main();
            PH(9);
            const double delta = ds.bc.delta;
            if (off > first) q_pos = fma(-delta, gc, q_pos);
            if (n_pos >= 0) { replay_next(); q_nxt = fma(-delta, gn, q_nxt); }
            ++t_chg;
            // (the ring entry was written by its owner before the barrier; single SM: fenced there too.  Cluster: the
            //  control lane's fence after the barrier orders it before ring_t)
            if (!CLUSTER && tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ring_t, t_chg); }
            pos0 += first + 1;
            while (ia < na && ds.srt_pos[ia] < pos0) ++ia;
            pa = ia < na ? ds.srt_j[ia] : -1;
            if (off <= first && pos < p) {                    // this thread's position is behind the window now
                if (n_pos >= 0) { advance(true); init_next(c_pos + GB_BLOCK); } else c_pos = -1;
            }
            // the entries of the predicted next change, one step ahead
            if (pa >= 0) {
                if (c_pos >= 0 && pgc_j != pa) { pg_c = gram_entry<DESIGN>(P, pa, c_j); pgc_j = pa; }
                if (n_pos >= 0 && pgn_j != pa) { pg_n = gram_entry<DESIGN>(P, pa, n_j); pgn_j = pa; }
            }
            PH(10);
        }
        SyncD::sync();
        sweep_epilogue<SyncD>(P, &sh, mask_s, beta_g, ds.scratch, key, c, s, sh.rr, nullptr, false);
        SyncD::sync();
    }
    PH(0);
    drain();
    if (ctl_lane) { ws_fence<CLUSTER>(); st_volatile_int(&peer->quit, 1); }
    if (!worker) return;
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2;
        P.hstate[GB_HS * c + 2] = sh.p0; P.hstate[GB_HS * c + 3] = sh.rr;
        P.hstate[GB_HS * c + 4] = __ldcg(P.hstate + GB_HS * c + 4) + (double)n_changes;
        P.hstate[GB_HS * c + 5] = __ldcg(P.hstate + GB_HS * c + 5) + (double)n_windows;
    }
#ifdef BLIP_PHASE_CLOCKS
    if (tid == 0 && c == 0)
        printf("ws2 decide | epilogue %lld wait-pre %lld sweepstart %lld fill0 %lld | eval %lld sync1 %lld nullfill %lld throttle %lld "
               "owner/gather %lld sync2 %lld apply+refill %lld\n", ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5],
               ph_t[6], ph_t[7], ph_t[8], ph_t[9], ph_t[10]);
#endif
}

template <int DESIGN, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? WS2_D_THREADS : WS_THREADS, 1) gibbs_gram_ws2_kernel(SweepParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double sm[];
    __shared__ WsCtl ctl;
    __shared__ Ws2DecideShared ds;

    const int p2 = P.p2;
    double* qbuf = sm;                                // [2][p2]  (cluster placement: one pair per CTA, kept equal by the helpers)
    uint32_t* mask_s = (uint32_t*)(sm + 2 * (size_t)p2);
    int c, role, rtid;                                // role 0 = decision (+ control warp), 1 = helper
    WsCtl* peer = &ctl;
    double* q_other = nullptr;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        role = (int)cluster.block_rank();
        c = blockIdx.x >> 1;
        rtid = threadIdx.x;
        peer = cluster.map_shared_rank(&ctl, role ^ 1);
        q_other = cluster.map_shared_rank(qbuf, role ^ 1);
    } else {
        role = threadIdx.x >= GB_BLOCK;
        c = blockIdx.x;
        rtid = threadIdx.x & (GB_BLOCK - 1);
    }
    {
        const double* q_g = P.q + (int64_t)c * p2;
        for (int k = threadIdx.x; k < p2; k += blockDim.x) qbuf[k] = __ldcg(q_g + k);
        for (int w = threadIdx.x; w < P.words; w += blockDim.x) mask_s[w] = __ldcg(P.mask + (int64_t)c * P.words + w);
    }
    if (threadIdx.x == 0) {
        ds.sh.sigma2 = __ldcg(P.hstate + GB_HS * c + 0); ds.sh.tau2 = __ldcg(P.hstate + GB_HS * c + 1);
        ds.sh.p0 = __ldcg(P.hstate + GB_HS * c + 2); ds.sh.rr = __ldcg(P.hstate + GB_HS * c + 3);
        ctl.pub = 0ull; ctl.ring_t = 0; ctl.ack_epoch = 0; ctl.d_sweep = 0; ctl.pre_ready = 0; ctl.quit = 0;
        ctl.f_act = 0; ctl.f_n = 0;
        ds.s_first[0] = ds.s_first[1] = ds.s_first[2] = GB_BLOCK;
    }
    if (CLUSTER) cg::this_cluster().sync();
    else __syncthreads();

    if (role == 1) { if (rtid < GB_BLOCK) ws_helper_role<DESIGN, CLUSTER>(P, c, rtid, qbuf, &ctl, peer, q_other); }
    else ws2_decision_role<DESIGN, CLUSTER>(P, c, qbuf, mask_s, &ctl, peer, ds, q_other);

    // a CTA's shared memory must outlive the peer's last access to it
    if (CLUSTER) cg::this_cluster().sync();
}
