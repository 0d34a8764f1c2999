// gibbs_gram_batch.cuh -- gram-form sweep for few chains, BATCHED decisions: the sweep as a
// bandwidth-bound column stream with a short sequential resolver instead of one latency chain.
//
// In the stationary regime nearly every change of a sweep is the redraw of a coordinate that was
// already active when the sweep started (K ~ 61 changes per sweep at C2, ~ 50 of them such redraws),
// and WHERE those happen is known at sweep start: the positions of the active coordinates in the
// visiting order.  What such a redraw needs is q at the active coordinates only, and a change at
// coordinate a moves q at coordinate b by delta * G[a][b]: a chain over |A|^2 Gram entries, not over
// Gram columns.  So:
//
//   resolver (one warp)   tracks q at up to WB_NA active coordinates in registers (two per lane) and
//                         resolves the next WB_B of them in visiting order -- decision, slab draw,
//                         delta, ||r||^2 -- ~350 cycles each, with the |A| Gram entries of a step
//                         fetched for the whole batch up front;
//   evaluators (8 warps)  check, in parallel and with all deltas known, every other position up to
//                         the batch's last one: q of the position's coordinate from the published
//                         buffer plus one gathered Gram entry per change it does not contain yet (the
//                         ring entries the helpers have not applied + the batch's changes before that
//                         position, <= WB_G in all) -- all loads independent, ONE L2 round trip;
//   helper warps          gibbs_gram_ws.cuh's helper role, unchanged: stream the Gram columns of the
//                         published changes into q (double-buffered) and precompute the next sweep.
//
// If no other position changes, the whole batch is committed: WB_B changes and several hundred null
// visits per three barriers.  If one does (an activation, ~5 per sweep, or an active coordinate beyond
// the WB_NA tracked ones), the batch's changes before it are committed, its owner applies it exactly
// like gibbs_gram_kernel's owner thread, the resolver rewinds to the state after the committed prefix
// (kept per step in registers) and gathers the new change's Gram entries.  From beta = 0 nothing is
// tracked and every change is found this way: one change per iteration, like the classic window.
//
// Every decision is the exact one (same guarded fast test, same fall-back to the reference formula),
// every coefficient is drawn by the same expression, and every element of q sees the same fma
// sequence in change order as in gibbs_gram_kernel, so results are bit-identical to it
// (tests/test_batch_gpu.py); only the work counter n_windows means something else here (iterations).
//
// Reference semantics: /root/reference/pyblip/linear/_linear.pyx:126-229 (unchanged).
#pragma once
#include "gibbs_gram_ws.cuh"

#ifndef BLIP_WB_B
#define BLIP_WB_B 4
#endif
#ifndef BLIP_WB_G
#define BLIP_WB_G 8
#endif
#ifndef BLIP_WB_VT
#define BLIP_WB_VT 3
#endif
constexpr int WB_B = BLIP_WB_B;            // tracked redraws resolved per batch
constexpr int WB_G = BLIP_WB_G;            // most gathered Gram entries per evaluated position (pending + batch)
constexpr int WB_VT = BLIP_WB_VT;          // positions a thread evaluates per iteration
constexpr int WB_V = WB_VT * GB_BLOCK;     // positions per iteration
constexpr int WB_NA = 64;                  // tracked active coordinates (two per resolver lane)
constexpr int WB_NONE = 0x7fffffff;
#ifdef BLIP_PHASE_CLOCKS
#define WPH_DECL long long ph_t[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0}; long long ph_last = clock64();
#define WPH(i) do { if (threadIdx.x == 0) { long long _n = clock64(); ph_t[i] += _n - ph_last; ph_last = _n; } } while (0)
#define WPH_WAIT(x) do { if ((x) == 1.2345e300) ph_t[15] += 1; } while (0)
#else
#define WPH_DECL
#define WPH(i) do {} while (0)
#define WPH_WAIT(x) do {} while (0)
#endif
static_assert(WB_G >= WB_B && WB_G + 1 <= WS_RING, "ring too small for the batch");

struct WbShared {
    SweepShared sh;
    int s_first[2];
    int act_n, na;
    int act_pos[GB_BLOCK];
    // tracked active coordinates of the sweep, in visiting order
    int srt_pos[WB_NA], srt_j[WB_NA];
    float srt_fA[WB_NA], srt_fc1[WB_NA], srt_fL[WB_NA];
    double srt_bold[WB_NA], srt_xl2[WB_NA], srt_zpx[WB_NA], srt_zpy[WB_NA];
    // the batch being decided
    int t_n, t_e1;
    int t_pos[WB_B], t_j[WB_B];
    double t_delta[WB_B], t_bnew[WB_B], t_rr[WB_B];
    // a change found by the evaluators
    int u_j; double u_delta, u_rr;
    double scratch[40];
};

// u <= kappa(XjTr): the guarded fast test of gibbs_gram_kernel, falling back to the reference formula
__device__ __forceinline__ bool wb_spike(double XjTr, double xl2, float fA, float fc1, float fL, const Hyper& h,
                                         const double* u_s, const RngKey& key, int pos, int sweep)
{
    const double x2c = (double)fc1 * (XjTr * XjTr);
    const double E = (double)fA - x2c;
    const double margin = E - (double)fL;
    const double guard = 1e-4 * (1.0 + fabs((double)fA) + x2c + fabs((double)fL));
    if (!(fabs(margin) > guard) || !(fmax(E, (double)fL) < 20.0)) {
        const double c_u = u_s ? u_s[pos] : rng_uniform(key, KIND_SPIKE_U, pos, sweep, 0);
        return c_u <= kappa_of(XjTr, xl2, h);
    }
    return margin > 0.0;
}

template <int DESIGN, bool CLUSTER>
__device__ __forceinline__ void wb_decision_role(const SweepParams& P, int c, double* q_d, uint32_t* mask_s, uint16_t* perm_sm,
                                                 WsCtl* mine, WsCtl* peer, WbShared& ds)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = P.p, p2 = P.p2, s_end = P.S;
    const RngKey key = { P.k0, P.k1, P.chain_offset + (uint32_t)c };
    SweepShared& sh = ds.sh;
    double* beta_g = P.beta + (int64_t)c * p;
    int eparity = 0;
    int epoch = 0, applied = 0, t_chg = 0;           // block-uniform
    auto adopt = [&]() {
        if (tid == 0) { const unsigned long long v = ld_volatile_u64(&mine->pub); ws_fence<CLUSTER>(); mine->slot[eparity] = v; }
        SyncD::sync();
        const unsigned long long v = mine->slot[eparity];
        eparity = eparity == 2 ? 0 : eparity + 1;
        epoch = (int)(v >> 32); applied = (int)(unsigned)v;
        if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ack_epoch, epoch); }
    };
    auto drain = [&]() { while (applied != t_chg) adopt(); };
    auto set_mask = [&](int j, bool on) {
        if (on) atomicOr(&mask_s[j >> 5], 1u << (j & 31)); else atomicAnd(&mask_s[j >> 5], ~(1u << (j & 31)));
    };
    auto publish_change = [&](int slot, int j, double delta, double bnew) {
        mine->ring_j[slot] = j; mine->ring_d[slot] = delta;
        if (CLUSTER) { peer->ring_j[slot] = j; peer->ring_d[slot] = delta; }
        __stcg(beta_g + j, bnew);
        set_mask(j, bnew != 0);
#ifndef BLIP_WB_NOPF
        if (DESIGN == BLIPGPU_DESIGN_DENSE)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.G + (int64_t)j * P.ldg),
                         "r"((unsigned)(p2 * sizeof(double))) : "memory");
#endif
        ws_fence<CLUSTER>();
    };

    long long n_changes = 0, n_iters = 0;
    int it = 0;
    WPH_DECL
    for (int s = 0; s < s_end; ++s) {
        WPH(0);
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
        WPH(1);
        // ---- periodic rebuild of q and ||r||^2 from q0 = X^T y (as in the other gram kernels)
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
            ws_fence<CLUSTER>();
        }
        if (tid == 0) { sh.logodds = log(sh.p0) - log(1 - sh.p0); ds.act_n = 0; }
        SyncD::sync();
        const Hyper h = { sh.sigma2, sh.tau2, sh.logodds };
        const double ts = h.tau2 / h.sigma2;
        const float f_logodds = (float)h.logodds, f_tau2 = (float)h.tau2, f_sigma2 = (float)h.sigma2;
        const float f_ts = (float)ts;
        const float* Ls = P.scrL + half;
        double2* Zs = P.scrZ + (int64_t)c * p;
        double rr = sh.rr;

        // ---- visiting order into shared memory; positions of the active coordinates
        for (int pos = tid; pos < p; pos += GB_BLOCK) {
            const int j = perm_s[pos];
            perm_sm[pos] = (uint16_t)j;
            if ((mask_s[j >> 5] >> (j & 31)) & 1u) {
                const int slot = atomicAdd(&ds.act_n, 1);
                if (slot < GB_BLOCK) ds.act_pos[slot] = pos;
                else slab_prepare(P, Zs, perm_s, h, key, z_s, pos, sweep);      // list full (early burn-in)
            }
        }
        SyncD::sync();
        const int n_list = min(ds.act_n, GB_BLOCK);
        if (tid < n_list) {
            // slab_prepare, and the first WB_NA in visiting order become the resolver's tracked set
            const int pos = ds.act_pos[tid], j = perm_sm[pos];
            const double zn = z_s ? z_s[pos] : rng_normal(key, KIND_SLAB_Z, pos, sweep, 0);
            const double xl2 = P.Xl2[j];
            const double pv = 1.0 / (1.0 / h.tau2 + xl2 / h.sigma2);
            const double2 zp = make_double2(sqrt(pv) * zn, pv);
            Zs[pos] = zp;
            int rank = 0;
            for (int k = 0; k < n_list; ++k) rank += ds.act_pos[k] < pos;
            if (rank < WB_NA) {
                ds.srt_pos[rank] = pos; ds.srt_j[rank] = j;
                ds.srt_bold[rank] = __ldcg(beta_g + j); ds.srt_xl2[rank] = xl2;
                ds.srt_zpx[rank] = zp.x; ds.srt_zpy[rank] = zp.y;
                ds.srt_fL[rank] = __ldcg(Ls + pos);
                const float xl2f = (float)xl2;
                ds.srt_fA[rank] = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
                ds.srt_fc1[rank] = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
            }
        }
        if (tid == 0) ds.na = min(n_list, WB_NA);
        adopt();                                             // (also the barrier after the block above)
        drain();                                             // every earlier change is in the published buffer
        const int na = ds.na;
        WPH(2);

        // resolver state (warp 0): q at the tracked coordinates, every change so far applied
        int mj0 = -1, mj1 = -1;
        double qa0 = 0.0, qa1 = 0.0, qs0 = 0.0, qs1 = 0.0;
        double qh0[WB_B], qh1[WB_B];
#pragma unroll
        for (int k = 0; k < WB_B; ++k) { qh0[k] = 0.0; qh1[k] = 0.0; }
        if (warp == 0) {
            if (lane < na) { mj0 = ds.srt_j[lane]; qa0 = q_d[(size_t)(epoch & 1) * p2 + mj0]; }
            if (lane + 32 < na) { mj1 = ds.srt_j[lane + 32]; qa1 = q_d[(size_t)(epoch & 1) * p2 + mj1]; }
        }

        int pos0 = 0, ia = 0;
        while (pos0 < p) {
            ++n_iters;
            while (t_chg - applied > WB_G - WB_B) adopt();   // room for a whole batch among the gathers
            const int npend = t_chg - applied;
            WPH(3);
            // ---- 1. the resolver decides the next tracked redraws, assuming nothing else changes
            if (warp == 0) {
                int e1 = min(pos0 + WB_V - 1, p - 1);
                int n = 0;
#pragma unroll
                for (int k = 0; k < WB_B; ++k) {
                    const int i = ia + k;
                    if (n == k && i < na && ds.srt_pos[i] <= e1) n = k + 1;
                }
                if (ia + n < na) e1 = min(e1, ds.srt_pos[ia + n] - 1);
                double g0[WB_B], g1[WB_B];
#pragma unroll
                for (int k = 0; k < WB_B; ++k) {
                    g0[k] = 0.0; g1[k] = 0.0;
                    if (k < n) {
                        const int jk = ds.srt_j[ia + k];
                        if (mj0 >= 0) g0[k] = gram_entry<DESIGN>(P, jk, mj0);
                        if (mj1 >= 0) g1[k] = gram_entry<DESIGN>(P, jk, mj1);
                    }
                }
                qs0 = qa0; qs1 = qa1;
                double rr_t = rr;
#ifdef BLIP_PHASE_CLOCKS
                { double sacc = 0.0;
#pragma unroll
                  for (int k = 0; k < WB_B; ++k) sacc += g0[k] + g1[k];
                  WPH_WAIT(sacc); }
                WPH(11);
#endif
#pragma unroll
                for (int k = 0; k < WB_B; ++k) {
                    if (k < n) {
                        const int i = ia + k;
                        const double q = __shfl_sync(0xffffffffu, (i & 32) ? qa1 : qa0, i & 31);
                        const double bold = ds.srt_bold[i], xl2 = ds.srt_xl2[i];
                        const int pos = ds.srt_pos[i];
                        const double XjTr = q + bold * xl2;
                        const bool spike = wb_spike(XjTr, xl2, ds.srt_fA[i], ds.srt_fc1[i], ds.srt_fL[i], h, u_s, key, pos, sweep);
                        double bnew = 0.0;
                        if (!spike) bnew = ds.srt_zpx[i] + ds.srt_zpy[i] * XjTr / h.sigma2;
                        const double delta = bnew - bold;
                        rr_t = rr_t - 2.0 * delta * q + delta * delta * xl2;
                        if (lane == 0) {
                            ds.t_pos[k] = pos; ds.t_j[k] = ds.srt_j[i];
                            ds.t_delta[k] = delta; ds.t_bnew[k] = bnew; ds.t_rr[k] = rr_t;
                        }
                        if (mj0 >= 0) qa0 = fma(-delta, g0[k], qa0);
                        if (mj1 >= 0) qa1 = fma(-delta, g1[k], qa1);
                        qh0[k] = qa0; qh1[k] = qa1;
                    }
                }
                if (lane == 0) { ds.t_n = n; ds.t_e1 = e1; }
            }
            WPH(4);
            SyncD::sync();
            WPH(5);
            const int n = ds.t_n, e1 = ds.t_e1;

            // ---- 2. everybody evaluates the other positions up to e1 with all deltas known
            int my_first = WB_NONE, o_j = 0;
            double o_q = 0.0, o_XjTr = 0.0, o_xl2 = 0.0, o_bold = 0.0;
            bool o_spike = true, o_act = false;
            {
                double dk[WB_G];
                int jk[WB_G];
#pragma unroll
                for (int k = 0; k < WB_G; ++k) {
                    dk[k] = 0.0; jk[k] = 0;
                    if (k < npend) {
                        const int slot = (applied + k) & (WS_RING - 1);
                        jk[k] = mine->ring_j[slot]; dk[k] = mine->ring_d[slot];
                    } else if (k - npend < n) {
                        jk[k] = ds.t_j[k - npend]; dk[k] = ds.t_delta[k - npend];
                    }
                }
                int jv[WB_VT], cnt[WB_VT];
                double qv[WB_VT], xv[WB_VT], gv[WB_VT][WB_G];
                float Lv[WB_VT];
                const double* qbuf = q_d + (size_t)(epoch & 1) * p2;
#pragma unroll
                for (int i = 0; i < WB_VT; ++i) {
                    const int pos = pos0 + tid + i * GB_BLOCK;
                    cnt[i] = -1; jv[i] = 0; qv[i] = 0.0; xv[i] = 0.0; Lv[i] = 0.f;
                    if (pos <= e1) {
                        const int j = perm_sm[pos];
                        int ct = npend;
                        bool own = false;
#pragma unroll
                        for (int k = 0; k < WB_B; ++k)
                            if (k < n) { const int tp = ds.t_pos[k]; ct += tp < pos; own |= tp == pos; }
                        if (!own) {
                            cnt[i] = ct; jv[i] = j;
                            qv[i] = qbuf[j]; xv[i] = P.Xl2[j]; Lv[i] = __ldcg(Ls + pos);
#pragma unroll
                            for (int k = 0; k < WB_G; ++k) gv[i][k] = k < ct ? gram_entry<DESIGN>(P, jk[k], j) : 0.0;
                        }
                    }
                }
#ifdef BLIP_PHASE_CLOCKS
                WPH(12);
                { double sacc = 0.0;
#pragma unroll
                  for (int i = 0; i < WB_VT; ++i) if (cnt[i] >= 0) { sacc += qv[i] + xv[i] + (double)Lv[i];
#pragma unroll
                      for (int k = 0; k < WB_G; ++k) sacc += gv[i][k]; }
                  WPH_WAIT(sacc); }
                WPH(13);
#endif
#pragma unroll
                for (int i = 0; i < WB_VT; ++i) {
                    if (cnt[i] >= 0) {
                        const int pos = pos0 + tid + i * GB_BLOCK, j = jv[i];
                        double q = qv[i];
#pragma unroll
                        for (int k = 0; k < WB_G; ++k) if (k < cnt[i]) q = fma(-dk[k], gv[i][k], q);
                        const bool act = (mask_s[j >> 5] >> (j & 31)) & 1u;
                        const double bold = act ? __ldcg(beta_g + j) : 0.0;
                        const double xl2 = xv[i];
                        const double XjTr = act ? q + bold * xl2 : q;
                        const float xl2f = (float)xl2;
                        const float fA = f_logodds + 0.5f * __logf(1.0f + __fdividef(f_tau2 * xl2f, f_sigma2));
                        const float fc1 = __fdividef(f_ts, 2.0f * (f_sigma2 + f_tau2 * xl2f));
                        const bool spike = wb_spike(XjTr, xl2, fA, fc1, Lv[i], h, u_s, key, pos, sweep);
                        if ((act || !spike) && pos < my_first) {
                            my_first = pos; o_j = j; o_q = q; o_XjTr = XjTr; o_xl2 = xl2; o_bold = bold;
                            o_spike = spike; o_act = act;
                        }
                    }
                }
            }
            {
                const int wmin = __reduce_min_sync(0xffffffffu, my_first);
                if (lane == 0 && wmin != WB_NONE) atomicMin(&ds.s_first[it & 1], wmin);
            }
            WPH(6);
            adopt();
            WPH(7);
            const int first = ds.s_first[it & 1];
            if (tid == 0) ds.s_first[(it + 1) & 1] = WB_NONE;
            ++it;
            const bool fail = first != WB_NONE;

            // ---- 3. commit the batch's changes before `first`, and the change at `first` itself
            int m = 0;
#pragma unroll
            for (int k = 0; k < WB_B; ++k) if (k < n && ds.t_pos[k] < first) m = k + 1;
            const double rr_m = m > 0 ? ds.t_rr[m - 1] : rr;  // (read before the barrier: the resolver rewrites t_* right after it)
            if (tid < m) publish_change((t_chg + tid) & (WS_RING - 1), ds.t_j[tid], ds.t_delta[tid], ds.t_bnew[tid]);
            if (fail && my_first == first) {                  // the owner of the changing position
                double bnew = 0.0;
                if (!o_spike) {
                    if (o_act) { const double2 zp = __ldcg(Zs + first); bnew = zp.x + zp.y * o_XjTr / h.sigma2; }
                    else bnew = slab_draw(o_XjTr, o_xl2, h, z_s ? z_s[first] : rng_normal(key, KIND_SLAB_Z, first, sweep, 0));
                }
                const double delta = bnew - o_bold;
                ds.u_rr = rr_m - 2.0 * delta * o_q + delta * delta * o_xl2;
                ds.u_delta = delta; ds.u_j = o_j;
                publish_change((t_chg + m) & (WS_RING - 1), o_j, delta, bnew);
            }
            WPH(8);
            SyncD::sync();
            WPH(9);
            rr = fail ? ds.u_rr : rr_m;
            t_chg += m + (fail ? 1 : 0);
            n_changes += m + (fail ? 1 : 0);
            if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->ring_t, t_chg); }
            if (fail) {
                if (warp == 0) {
                    // back to the state after the m committed steps, then the change nobody had planned
                    if (m == 0) { qa0 = qs0; qa1 = qs1; }
#pragma unroll
                    for (int k = 0; k < WB_B; ++k) if (k == m - 1) { qa0 = qh0[k]; qa1 = qh1[k]; }
                    const int ju = ds.u_j;
                    const double du = ds.u_delta;
                    if (mj0 >= 0) qa0 = fma(-du, gram_entry<DESIGN>(P, ju, mj0), qa0);
                    if (mj1 >= 0) qa1 = fma(-du, gram_entry<DESIGN>(P, ju, mj1), qa1);
                }
                ia += m; pos0 = first + 1;
            } else {
                ia += n; pos0 = e1 + 1;
            }
            WPH(10);
        }
        if (tid == 0) sh.rr = rr;
        SyncD::sync();
        sweep_epilogue<SyncD>(P, &sh, mask_s, beta_g, ds.scratch, key, c, s, rr, nullptr, false);
        SyncD::sync();
    }
    WPH(0);
    // ---- launch end: everything into the published buffer (the helpers write it back to HBM)
    drain();
    if (tid == 0) { ws_fence<CLUSTER>(); st_volatile_int(&peer->quit, 1); }
    for (int w = tid; w < P.words; w += GB_BLOCK) P.mask[(int64_t)c * P.words + w] = mask_s[w];
    if (tid == 0) {
        P.hstate[GB_HS * c + 0] = sh.sigma2; P.hstate[GB_HS * c + 1] = sh.tau2;
        P.hstate[GB_HS * c + 2] = sh.p0; P.hstate[GB_HS * c + 3] = sh.rr;
        P.hstate[GB_HS * c + 4] = __ldcg(P.hstate + GB_HS * c + 4) + (double)n_changes;
        P.hstate[GB_HS * c + 5] = __ldcg(P.hstate + GB_HS * c + 5) + (double)n_iters;
    }
#ifdef BLIP_PHASE_CLOCKS
    if (tid == 0 && c == 0)
        printf("wb decide | epilogue %lld wait-pre %lld sweepstart %lld | throttle %lld resolve: issue+wait %lld steps %lld syncA %lld evaluate: issue %lld wait %lld math %lld "
               "syncB %lld commit %lld syncC %lld rewind %lld | iters %lld changes %lld\n", ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[11], ph_t[4], ph_t[5],
               ph_t[12], ph_t[13], ph_t[6], ph_t[7], ph_t[8], ph_t[9], ph_t[10], n_iters, n_changes);
#endif
}

template <int DESIGN, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? GB_BLOCK : WS_THREADS, 1) gibbs_gram_batch_kernel(SweepParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double sm[];
    __shared__ WsCtl ctl;
    __shared__ WbShared ds;

    const int p2 = P.p2;
    double* qbuf = sm;                                // [2][p2]  (cluster placement: used in the helper CTA only)
    uint32_t* mask_s = (uint32_t*)(sm + 2 * (size_t)p2);
    uint16_t* perm_sm = (uint16_t*)(mask_s + P.words);
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
        ds.s_first[0] = ds.s_first[1] = WB_NONE;
    }
    if (CLUSTER) cg::this_cluster().sync();
    else __syncthreads();

    if (role == 1) ws_helper_role<DESIGN, CLUSTER>(P, c, rtid, qbuf, &ctl, peer);
    else wb_decision_role<DESIGN, CLUSTER>(P, c, q_d, mask_s, perm_sm, &ctl, peer, ds);

    if (CLUSTER) cg::this_cluster().sync();
}
