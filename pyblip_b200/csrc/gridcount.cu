// gridcount.cu -- continuous-space candidate regions: per-sample de-duplicated cell /
// circle membership counts.
//
// Replaces the Python loops of /root/reference/pyblip/create_groups_cts.py:152-206
// (`grid_peps`: per sample, per grid size, bucket every signal location into its grid cell --
// and, for circles, the four axis-neighbour circles that also contain it -- de-duplicate
// per sample, accumulate key -> count).  Keys stay integers on the device:
//     code = [extra:1][grid index:7][cell coordinates + 1 : 56 bits split over d <= 3 dims]
// and the host turns a code back into the reference's float key with the reference's own
// expression (floor(x*g)/g + 1/(2g), rounded to 10 decimals), so keys are bit-identical.
//
// One CTA per (sample, grid size): codes go into a shared-memory hash set (per-sample
// de-duplication, with multiplicities), every distinct code is then counted once in a
// global open-addressing table (64-bit CAS + 32-bit atomic add).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>
#include "internal.cuh"

struct blipgpu_gridcounts {
    blipgpu_ctx* ctx;
    int64_t n_keys, n_pairs;
    unsigned long long* codes; unsigned int* counts;       // device, compacted
    unsigned long long* pair_codes; unsigned int* pair_mult; // device, per (sample, code) multiplicities
};

namespace {

constexpr unsigned long long EMPTY = ~0ull;
constexpr int GC_BLOCK = 128;

struct GCParams {
    const double* locs; int64_t N; int n_src, d;
    const double* grid; int G; int circle;
    int bits_per_dim; int smem_cap;            // slots of the per-block hash set (power of two)
    unsigned long long* tab_keys; unsigned int* tab_cnt; unsigned long long tab_mask;
    unsigned long long* pair_codes; unsigned int* pair_mult; unsigned long long* n_pairs; long long pair_cap;
    int want_pairs; int* err;
    // extra centers: code to count for (centre, grid) -- its own code, or the code of the grid
    // cell whose key it coincides with (the reference merges equal keys per sample)
    const double* extra; int n_extra; const double* radii; const unsigned long long* extra_codes;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ void global_count(const GCParams& P, unsigned long long code)
{
    unsigned long long h = mix64(code) & P.tab_mask;
    for (unsigned long long probe = 0; probe <= P.tab_mask; ++probe) {
        unsigned long long prev = atomicCAS(P.tab_keys + h, EMPTY, code);
        if (prev == EMPTY || prev == code) { atomicAdd(P.tab_cnt + h, 1u); return; }
        h = (h + 1) & P.tab_mask;
    }
    *P.err = 1;    // table full
}

__device__ __forceinline__ void emit_pair(const GCParams& P, unsigned long long code, unsigned int mult)
{
    if (!P.want_pairs) return;
    unsigned long long at = atomicAdd(P.n_pairs, 1ull);
    if ((long long)at < P.pair_cap) { P.pair_codes[at] = code; P.pair_mult[at] = mult; }
    else *P.err = 2;
}

__device__ __forceinline__ void smem_insert(unsigned long long* keys, unsigned int* cnt, int cap, unsigned long long code,
                                            unsigned int times = 1u)
{
    unsigned int h = (unsigned int)mix64(code) & (cap - 1);
    for (int probe = 0; probe < cap; ++probe) {
        unsigned long long prev = atomicCAS(keys + h, EMPTY, code);
        if (prev == EMPTY || prev == code) { atomicAdd(cnt + h, times); return; }
        h = (h + 1) & (cap - 1);
    }
}

// grid (G, N): block = (sample, grid size)
__global__ void __launch_bounds__(GC_BLOCK) grid_codes_kernel(GCParams P)
{
    extern __shared__ unsigned long long sm_keys[];
    unsigned int* sm_cnt = (unsigned int*)(sm_keys + P.smem_cap);
    const int gi = blockIdx.x;
    const int64_t s = blockIdx.y;
    const double g = P.grid[gi];
    const int d = P.d, bits = P.bits_per_dim;
    for (int k = threadIdx.x; k < P.smem_cap; k += GC_BLOCK) { sm_keys[k] = EMPTY; sm_cnt[k] = 0u; }
    __syncthreads();
    const unsigned long long head = (unsigned long long)gi << 56;
    for (int i = threadIdx.x; i < P.n_src; i += GC_BLOCK) {
        const double* x = P.locs + ((int64_t)s * P.n_src + i) * d;
        bool ok = true;
        double xs[3] = { 0.0, 0.0, 0.0 };
        long long cell[3] = { 0, 0, 0 };
        for (int k = 0; k < d; ++k) { xs[k] = x[k]; ok = ok && !isnan(xs[k]); }
        if (!ok) continue;                                   // dummy discovery (:154)
        unsigned long long code = head;
        for (int k = 0; k < d; ++k) {
            cell[k] = (long long)floor(__dmul_rn(xs[k], g));     // np.floor(samples * gsize)
            code |= (unsigned long long)(cell[k] + 1) << (bits * k);
        }
        smem_insert(sm_keys, sm_cnt, P.smem_cap, code);
        if (P.circle) {
            // _additional_circle_centers (:47-66), d == 2
            const double half = 1 / (2 * g);
            const double radius = sqrt((double)d) / (2 * g);
            const double radius2 = __dmul_rn(radius, radius);
            double cen[2];
            for (int k = 0; k < 2; ++k) cen[k] = __dadd_rn(__ddiv_rn((double)cell[k], g), half);
            for (int k = 0; k < 2; ++k)
                for (int off = -1; off <= 1; off += 2) {
                    double cn[2] = { cen[0], cen[1] };
                    cn[k] = __dadd_rn(cn[k], __ddiv_rn((double)off, g));
                    const double dx = __dadd_rn(xs[0], -cn[0]), dy = __dadd_rn(xs[1], -cn[1]);
                    const double dist2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    bool inc = dist2 <= radius2;
                    inc = inc && (fmax(cn[0], cn[1]) < 1 + 1e-10);
                    inc = inc && (fmin(cn[0], cn[1]) > -1e-10);
                    if (inc) {
                        long long c2[2] = { cell[0], cell[1] };
                        c2[k] += off;
                        unsigned long long nc = head | (unsigned long long)(c2[0] + 1) | ((unsigned long long)(c2[1] + 1) << bits);
                        smem_insert(sm_keys, sm_cnt, P.smem_cap, nc);
                    }
                }
        }
    }
    // extra centres (:165-184): a centre within `radius` of the nearest signal counts once
    // (or once per signal within the radius when multiplicities are wanted)
    for (int nc = threadIdx.x; nc < P.n_extra; nc += GC_BLOCK) {
        const double* c = P.extra + (int64_t)nc * d;
        const double radius = P.radii[gi];
        double mind = INFINITY;
        unsigned int nsig = 0;
        for (int i = 0; i < P.n_src; ++i) {
            const double* x = P.locs + ((int64_t)s * P.n_src + i) * d;
            bool ok = true;
            for (int k = 0; k < d; ++k) ok = ok && !isnan(x[k]);
            if (!ok) continue;
            double dist = 0.0;
            if (P.circle) {
                for (int k = 0; k < d; ++k) { const double df = __dadd_rn(c[k], -x[k]); dist = __dadd_rn(dist, __dmul_rn(df, df)); }
                dist = sqrt(dist);
            } else {
                for (int k = 0; k < d; ++k) dist = fmax(dist, fabs(__dadd_rn(c[k], -x[k])));
            }
            mind = fmin(mind, dist);
            nsig += dist <= radius ? 1u : 0u;
        }
        if (mind <= radius)
            smem_insert(sm_keys, sm_cnt, P.smem_cap, P.extra_codes[(int64_t)nc * P.G + gi], P.want_pairs ? nsig : 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < P.smem_cap; k += GC_BLOCK) {
        const unsigned long long code = sm_keys[k];
        if (code != EMPTY) { global_count(P, code); emit_pair(P, code, sm_cnt[k]); }
    }
}

__global__ void compact_table_kernel(const unsigned long long* keys, const unsigned int* cnt, unsigned long long cap,
                                     unsigned long long* out_codes, unsigned int* out_cnt, unsigned long long* n_out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long k = keys[i];
    if (k != EMPTY) {
        const unsigned long long at = atomicAdd(n_out, 1ull);
        out_codes[at] = k; out_cnt[at] = cnt[i];
    }
}

} // namespace

extern "C" int blipgpu_grid_count(blipgpu_ctx* ctx, const double* locs, int64_t N, int64_t n_src, int32_t d,
                                  const double* grid_sizes, int32_t G, int32_t circle,
                                  const double* extra_centers, int64_t n_extra, const double* extra_radii,
                                  const uint64_t* extra_codes, int32_t want_pairs, blipgpu_gridcounts** out)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!locs || !grid_sizes || !out || N < 0 || n_src <= 0 || d <= 0 || G <= 0) { blip_set_error("grid_count: bad arguments"); return BLIP_ERR_ARG; }
    if (d > 3) { blip_set_error("grid_count: d = %d > 3 is not supported", d); return BLIP_ERR_UNSUPPORTED; }
    if (circle && d != 2) { blip_set_error("Shape=circle only implemented for 2-d data"); return BLIP_ERR_UNSUPPORTED; }
    if (G > 127) { blip_set_error("grid_count: at most 127 grid sizes"); return BLIP_ERR_UNSUPPORTED; }
    if (n_extra > 0 && (!extra_centers || !extra_radii || !extra_codes)) { blip_set_error("grid_count: extra centers without radii"); return BLIP_ERR_ARG; }
    const int bits = d == 1 ? 56 : d == 2 ? 28 : 18;
    for (int g = 0; g < G; ++g)
        if (!(grid_sizes[g] >= 1) || grid_sizes[g] + 3 >= (double)(1ull << bits)) { blip_set_error("grid_count: grid size %g out of range for d = %d", grid_sizes[g], d); return BLIP_ERR_ARG; }
    cudaStream_t st = ctx->stream;

    blipgpu_gridcounts* res = new (std::nothrow) blipgpu_gridcounts();
    if (!res) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    memset(res, 0, sizeof(*res));
    res->ctx = ctx;
    struct Guard { blipgpu_gridcounts* r; ~Guard() { if (r) blipgpu_gridcounts_free(r); } } guard{ res };

    const int per_point = circle ? 5 : 1;
    int cap = 64;
    while (cap < 2 * (n_src * per_point + n_extra)) cap <<= 1;
    const size_t smem = (size_t)cap * (sizeof(unsigned long long) + sizeof(unsigned int));
    if (smem > ctx->smem_optin) { blip_set_error("grid_count: %lld sources per sample exceed the shared-memory hash set", (long long)n_src); return BLIP_ERR_UNSUPPORTED; }
    const double emitted = (double)N * n_src * G * per_point + (double)N * n_extra * G;   // upper bound on codes
    unsigned long long tab = 1ull << 16;
    while ((double)tab < 2.0 * emitted && tab < (1ull << 27)) tab <<= 1;
    const long long pair_cap = want_pairs ? (long long)std::min(emitted, 4.0e8) + 1 : 1;

    double *d_locs = nullptr, *d_grid = nullptr, *d_extra = nullptr, *d_radii = nullptr;
    unsigned long long* d_ecodes = nullptr;
    unsigned long long *d_keys = nullptr, *d_npairs = nullptr, *d_nout = nullptr;
    unsigned int* d_cnt = nullptr; int* d_err = nullptr;
    auto cleanup = [&]() { cudaFree(d_locs); cudaFree(d_grid); cudaFree(d_extra); cudaFree(d_radii); cudaFree(d_ecodes); cudaFree(d_keys); cudaFree(d_cnt); cudaFree(d_npairs); cudaFree(d_nout); cudaFree(d_err); };
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, std::max<size_t>(b, 8)); };
    A((void**)&d_locs, sizeof(double) * N * n_src * d); A((void**)&d_grid, sizeof(double) * G);
    A((void**)&d_extra, sizeof(double) * n_extra * d); A((void**)&d_radii, sizeof(double) * G);
    A((void**)&d_ecodes, sizeof(unsigned long long) * n_extra * G);
    A((void**)&d_keys, sizeof(unsigned long long) * tab); A((void**)&d_cnt, sizeof(unsigned int) * tab);
    A((void**)&d_npairs, 8); A((void**)&d_nout, 8); A((void**)&d_err, 4);
    A((void**)&res->pair_codes, sizeof(unsigned long long) * pair_cap); A((void**)&res->pair_mult, sizeof(unsigned int) * pair_cap);
    if (e != cudaSuccess) { cleanup(); blip_set_error("grid_count: allocation failed: %s", cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    if (N > 0) cudaMemcpyAsync(d_locs, locs, sizeof(double) * N * n_src * d, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_grid, grid_sizes, sizeof(double) * G, cudaMemcpyHostToDevice, st);
    if (n_extra > 0) {
        cudaMemcpyAsync(d_extra, extra_centers, sizeof(double) * n_extra * d, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_radii, extra_radii, sizeof(double) * G, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_ecodes, extra_codes, sizeof(unsigned long long) * n_extra * G, cudaMemcpyHostToDevice, st);
    }
    ctx->h2d_bytes += (int64_t)sizeof(double) * (N * n_src * d + G + n_extra * d);
    cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * tab, st);
    cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int) * tab, st);
    cudaMemsetAsync(d_npairs, 0, 8, st); cudaMemsetAsync(d_nout, 0, 8, st); cudaMemsetAsync(d_err, 0, 4, st);

    GCParams P;
    memset(&P, 0, sizeof(P));
    P.locs = d_locs; P.N = N; P.n_src = (int)n_src; P.d = d; P.grid = d_grid; P.G = G; P.circle = circle;
    P.bits_per_dim = bits; P.smem_cap = cap; P.tab_keys = d_keys; P.tab_cnt = d_cnt; P.tab_mask = tab - 1;
    P.pair_codes = res->pair_codes; P.pair_mult = res->pair_mult; P.n_pairs = d_npairs; P.pair_cap = pair_cap;
    P.want_pairs = want_pairs; P.err = d_err; P.extra = d_extra; P.n_extra = (int)n_extra; P.radii = d_radii; P.extra_codes = d_ecodes;
    if (N > 0) {
        ScopedTimer timer(ctx, false);
        cudaFuncSetAttribute(grid_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int launches = 0;
        for (int64_t s0 = 0; s0 < N; s0 += 65535) {           // grid.y limit
            GCParams Q = P;
            Q.locs = d_locs + s0 * n_src * d; Q.N = std::min<int64_t>(65535, N - s0);
            grid_codes_kernel<<<dim3((unsigned)G, (unsigned)Q.N), GC_BLOCK, smem, st>>>(Q);
            ++launches;
        }
        timer.stop(launches);
    }
    int err = 0; unsigned long long npairs = 0;
    cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&npairs, d_npairs, 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cleanup(); blip_set_error("grid_count: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    if (err) { cleanup(); blip_set_error(err == 1 ? "grid_count: key table full" : "grid_count: pair list full"); return BLIP_ERR_NOMEM; }
    // compact
    e = cudaMalloc(&res->codes, sizeof(unsigned long long) * tab);
    if (e == cudaSuccess) e = cudaMalloc(&res->counts, sizeof(unsigned int) * tab);
    if (e != cudaSuccess) { cleanup(); blip_set_error("grid_count: allocation failed: %s", cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    compact_table_kernel<<<(unsigned)((tab + 255) / 256), 256, 0, st>>>(d_keys, d_cnt, tab, res->codes, res->counts, d_nout);
    unsigned long long nout = 0;
    cudaMemcpyAsync(&nout, d_nout, 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    ctx->kernel_launches += 1;
    cleanup();
    if (e != cudaSuccess) { blip_set_error("grid_count: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    res->n_keys = (int64_t)nout; res->n_pairs = want_pairs ? (int64_t)npairs : 0;
    guard.r = nullptr;
    *out = res;
    return BLIP_OK;
}

extern "C" int blipgpu_gridcounts_sizes(blipgpu_gridcounts* r, int64_t* n_keys, int64_t* n_pairs)
{
    if (!r) { blip_set_error("gridcounts_sizes: NULL handle"); return BLIP_ERR_ARG; }
    if (n_keys) *n_keys = r->n_keys;
    if (n_pairs) *n_pairs = r->n_pairs;
    return BLIP_OK;
}

extern "C" int blipgpu_gridcounts_get(blipgpu_gridcounts* r, uint64_t* codes, uint32_t* counts,
                                      uint64_t* pair_codes, uint32_t* pair_mult)
{
    if (!r) { blip_set_error("gridcounts_get: NULL handle"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(r->ctx));
    if (codes && r->n_keys) CUDA_TRY(cudaMemcpy(codes, r->codes, sizeof(uint64_t) * r->n_keys, cudaMemcpyDeviceToHost));
    if (counts && r->n_keys) CUDA_TRY(cudaMemcpy(counts, r->counts, sizeof(uint32_t) * r->n_keys, cudaMemcpyDeviceToHost));
    if (pair_codes && r->n_pairs) CUDA_TRY(cudaMemcpy(pair_codes, r->pair_codes, sizeof(uint64_t) * r->n_pairs, cudaMemcpyDeviceToHost));
    if (pair_mult && r->n_pairs) CUDA_TRY(cudaMemcpy(pair_mult, r->pair_mult, sizeof(uint32_t) * r->n_pairs, cudaMemcpyDeviceToHost));
    r->ctx->d2h_bytes += 12 * (r->n_keys + r->n_pairs);
    return BLIP_OK;
}

extern "C" int blipgpu_gridcounts_free(blipgpu_gridcounts* r)
{
    if (!r) return BLIP_OK;
    cudaSetDevice(r->ctx->device);
    cudaFree(r->codes); cudaFree(r->counts); cudaFree(r->pair_codes); cudaFree(r->pair_mult);
    delete r;
    return BLIP_OK;
}


// ---------------------------------------------------------------------------
// Overlap (constraint) matrix of candidate regions: the O(G^2) step of
// /root/reference/pyblip/create_groups_cts.py:283-308, in the reference's float32 arithmetic
// (the file is compiled with -fmad=false, sqrtf is correctly rounded):
//   squares: all_j |c_a[j] - c_b[j]| < r_a + r_b        circles: sqrt(sum_j (c_a[j] - c_b[j])^2) < r_a + r_b
// Output: bit-packed rows, bit b of word [a][b / 32].  One warp per (row, 32-column word):
// lane = column, one ballot per word.
// ---------------------------------------------------------------------------
namespace {
__global__ void overlap_bits_kernel(const float* __restrict__ centers, const float* __restrict__ radii,
                                    int64_t G, int d, int circle, int words, uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= G * words) return;
    const int64_t a = gw / words;
    const int w = (int)(gw % words);
    const int64_t b = 32LL * w + lane;
    bool hit = false;
    if (b < G) {
        const float delta = radii[a] + radii[b];
        if (circle) {
            float ssum = 0.f;
            for (int j = 0; j < d; ++j) {
                const float diff = centers[(int64_t)j * G + a] - centers[(int64_t)j * G + b];
                const float sq = diff * diff;
                ssum = j == 0 ? sq : ssum + sq;
            }
            hit = sqrtf(ssum) < delta;
        } else {
            hit = true;
            for (int j = 0; j < d; ++j)
                hit = hit && (fabsf(centers[(int64_t)j * G + a] - centers[(int64_t)j * G + b]) < delta);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) out[a * words + w] = m;
}
} // namespace

extern "C" int blipgpu_overlap_bits(blipgpu_ctx* ctx, const float* centers, const float* radii, int64_t G,
                                    int32_t d, int32_t circle, uint32_t* out_bits)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!centers || !radii || !out_bits || G <= 0 || d <= 0) { blip_set_error("overlap_bits: bad arguments"); return BLIP_ERR_ARG; }
    if (G > 131072) { blip_set_error("overlap_bits: at most 131072 regions (the bit matrix would exceed 2 GB)"); return BLIP_ERR_UNSUPPORTED; }
    const int words = (int)((G + 31) / 32);
    float *d_c = nullptr, *d_r = nullptr; uint32_t* d_o = nullptr;
    cudaError_t e = cudaMalloc(&d_c, sizeof(float) * d * G);
    if (e == cudaSuccess) e = cudaMalloc(&d_r, sizeof(float) * G);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, sizeof(uint32_t) * G * words);
    cudaStream_t st = ctx->stream;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, centers, sizeof(float) * d * G, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_r, radii, sizeof(float) * G, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ScopedTimer timer(ctx, false);
        const int64_t nwarp = G * words;
        overlap_bits_kernel<<<(unsigned)((nwarp + 7) / 8), 256, 0, st>>>(d_c, d_r, G, d, circle, words, d_o);
        e = cudaGetLastError();
        timer.stop(1);
    }
    int rc = BLIP_OK;
    if (e == cudaSuccess) rc = blip_d2h(ctx, out_bits, d_o, sizeof(uint32_t) * G * words);
    if (e != cudaSuccess) { blip_set_error("overlap_bits: %s", cudaGetErrorString(e)); rc = BLIP_ERR_CUDA; }
    ctx->h2d_bytes += (int64_t)sizeof(float) * (d + 1) * G;
    ctx->d2h_bytes += (int64_t)sizeof(uint32_t) * G * words;
    cudaFree(d_c); cudaFree(d_r); cudaFree(d_o);
    return rc;
}
