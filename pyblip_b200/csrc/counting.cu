// counting.cu -- candidate-group posterior-inclusion counting on bit-packed samples.
//
// Replaces the array math of /root/reference/pyblip/create_groups.py:
//   :105-106  marg_pips = mean(samples != 0, axis=0)            -> count_columns
//   :321-332  cum_incs / cum_diffs == 0 / mean over rows          -> count_sequential
//   :459      1 - mean(any(samples[:, group], axis=1))            -> count_groups
//   :117      samples[:, rel_features]                            -> select_columns
// Everything is integer counting; the PEP is formed on the host as one correctly
// rounded divide, so results are bit-identical to the reference's float64 means.
//
// Layout: bits[w][row] (word-column major).  A warp owns one 32-column word w and
// walks rows with lane = row, so every load is a coalesced 128-byte line; per
// column totals come from a 32x32 in-warp bit transpose followed by popc.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>
#include "internal.cuh"
#include "pairs_tc.cuh"

namespace {

constexpr int CT_BLOCK = 256;
constexpr int CT_NWARP = CT_BLOCK / 32;

// Device buffers of the counting calls come from the device's stream-ordered pool on the context's stream, like the
// sampler's (internal.cuh): a plain cudaMalloc / cudaFree of the GB-sized ones (the column-major copy of the bit
// matrix, selected-column sets) was seen to take hundreds of ms now and then, and every cudaFree synchronises the device.
inline cudaError_t dev_alloc(blipgpu_ctx* ctx, void** p, size_t bytes) { return blip_store_alloc(p, bytes, ctx->stream); }
inline void dev_free(blipgpu_ctx* ctx, void* p) { blip_store_free(p, ctx->stream); }

// In-warp 32x32 bit-matrix transpose: on entry lane r holds row r (bit c = column c);
// on exit lane c holds column c (bit r = row r).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        // swap the off-diagonal s x s blocks between lane and lane^s
        const uint32_t mask = (s == 16) ? 0x0000FFFFu : (s == 8) ? 0x00FF00FFu : (s == 4) ? 0x0F0F0F0Fu
                            : (s == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        if (lane & s) x = (x & ~mask) | ((y >> s) & mask);
        else          x = (x & mask) | ((y << s) & ~mask);
    }
    return x;
}

// ---- dense -> bits -------------------------------------------------------
template <typename T>
__global__ void pack_dense_kernel(const T* __restrict__ dense, int64_t nrows, int64_t p, int words,
                                  int64_t row0, int64_t N_pad, uint32_t* __restrict__ bits)
{
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= nrows * words) return;
    const int64_t r = gw / words;
    const int w = (int)(gw % words);
    const int64_t col = 32LL * w + (threadIdx.x & 31);
    bool nz = false;
    if (col < p) nz = dense[r * p + col] != (T)0;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0) bits[(int64_t)w * N_pad + row0 + r] = m;
}

__global__ void clear_first_column_kernel(uint32_t* bits, int64_t N)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) bits[r] &= ~1u;
}

// ---- samples[:, cols] ----------------------------------------------------
__global__ void select_columns_kernel(const uint32_t* __restrict__ src, int64_t N, int64_t N_pad,
                                      const int32_t* __restrict__ cols, int ncols, int out_words,
                                      uint32_t* __restrict__ dst)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (r >= N) return;
    uint32_t out = 0u;
    int last_w = -1;
    uint32_t last = 0u;
    const int nb = min(32, ncols - 32 * w);
    for (int b = 0; b < nb; ++b) {
        const int cidx = cols[32 * w + b];
        const int sw = cidx >> 5;
        if (sw != last_w) { last = src[(int64_t)sw * N_pad + r]; last_w = sw; }
        out |= ((last >> (cidx & 31)) & 1u) << b;
    }
    dst[(int64_t)w * N_pad + r] = out;
}

__global__ void unpack_kernel(const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int64_t p,
                              int64_t row0, int64_t nrows, uint8_t* __restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * p) return;
    const int64_t r = idx / p, j = idx % p;
    out[idx] = (bits[(j >> 5) * N_pad + row0 + r] >> (j & 31)) & 1u;
}

// ---- column counts -------------------------------------------------------
// grid (words, chunks); each warp walks 32-row tiles of its chunk, lane = row.  A lane never
// transposes its words: it adds them into a bit-sliced ("vertical") counter -- plane k holds bit
// k of the 32 per-column counts -- with a carry-save adder tree over 16 words at a time
// (Harley-Seal) and a ripple into the upper planes, ~4 logic instructions per word, so the kernel
// streams.  Only the final planes are transposed (lane = column) and pop-counted.
#define CSA(h, l, a, b, c) do { const uint32_t _u = (a) ^ (b); h = ((a) & (b)) | (_u & (c)); l = _u ^ (c); } while (0)
constexpr int CC_PLANES = 20;      // upper planes: weights 16 .. 16 * 2^19 words per lane
__global__ void __launch_bounds__(CT_BLOCK) count_columns_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int64_t p, int64_t rows_per_block,
    unsigned long long* __restrict__ counts)
{
    __shared__ unsigned long long s_cnt[CT_NWARP][32];
    const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(N, r_begin + rows_per_block);
    const uint32_t* col = bits + (int64_t)w * N_pad;
    uint32_t ones = 0u, twos = 0u, fours = 0u, eights = 0u, hi[CC_PLANES];
#pragma unroll
    for (int k = 0; k < CC_PLANES; ++k) hi[k] = 0u;
    const int64_t step = 32 * CT_NWARP;
    int64_t r = r_begin + 32 * warp + lane;
    for (; r + 15 * step < r_end; r += 16 * step) {        // 16 live words
        uint32_t x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = col[r + i * step];
        uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteens;
        CSA(twosA, ones, ones, x[0], x[1]);    CSA(twosB, ones, ones, x[2], x[3]);
        CSA(foursA, twos, twos, twosA, twosB);
        CSA(twosA, ones, ones, x[4], x[5]);    CSA(twosB, ones, ones, x[6], x[7]);
        CSA(foursB, twos, twos, twosA, twosB);
        CSA(eightsA, fours, fours, foursA, foursB);
        CSA(twosA, ones, ones, x[8], x[9]);    CSA(twosB, ones, ones, x[10], x[11]);
        CSA(foursA, twos, twos, twosA, twosB);
        CSA(twosA, ones, ones, x[12], x[13]);  CSA(twosB, ones, ones, x[14], x[15]);
        CSA(foursB, twos, twos, twosA, twosB);
        CSA(eightsB, fours, fours, foursA, foursB);
        CSA(sixteens, eights, eights, eightsA, eightsB);
        uint32_t carry = sixteens;                         // ripple into the planes of weight >= 16
#pragma unroll
        for (int k = 0; k < CC_PLANES; ++k) { const uint32_t t = hi[k] & carry; hi[k] ^= carry; carry = t; }
    }
    for (; r < r_end; r += step) {                         // remaining words, one at a time
        uint32_t carry = col[r], t;
        t = ones & carry; ones ^= carry; carry = t;
        t = twos & carry; twos ^= carry; carry = t;
        t = fours & carry; fours ^= carry; carry = t;
        t = eights & carry; eights ^= carry; carry = t;
#pragma unroll
        for (int k = 0; k < CC_PLANES; ++k) { t = hi[k] & carry; hi[k] ^= carry; carry = t; }
    }
    // per-column totals of this warp: transpose every plane (lane c <- column c), pop-count over rows
    unsigned long long tot = (unsigned long long)__popc(warp_transpose32(ones)) +
                             2ull * __popc(warp_transpose32(twos)) + 4ull * __popc(warp_transpose32(fours)) +
                             8ull * __popc(warp_transpose32(eights));
#pragma unroll
    for (int k = 0; k < CC_PLANES; ++k)
        if (__any_sync(0xffffffffu, hi[k] != 0u)) tot += ((unsigned long long)__popc(warp_transpose32(hi[k]))) << (4 + k);
    s_cnt[warp][lane] = tot;
    __syncthreads();
    if (warp == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i][lane];
        const int64_t j = 32LL * w + lane;
        if (j < p && t) atomicAdd(counts + j, t);
    }
}

// ---- sequential (contiguous-window) zero counts --------------------------
// zero_counts[m][j] = #rows with bits j..j+m all zero, m < max_size, j + m < p.
template <int MAXM>
__global__ void __launch_bounds__(CT_BLOCK) count_sequential_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int64_t p, int words, int max_size,
    int64_t rows_per_block, unsigned long long* __restrict__ zero_counts)
{
    __shared__ int s_cnt[CT_NWARP][32];
    const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(N, r_begin + rows_per_block);
    const uint32_t* c0 = bits + (int64_t)w * N_pad;
    const uint32_t* c1 = (w + 1 < words) ? bits + (int64_t)(w + 1) * N_pad : nullptr;
    const uint32_t* c2 = (MAXM > 32 && w + 2 < words) ? bits + (int64_t)(w + 2) * N_pad : nullptr;

    // Column-major view of a 32-row tile: after the in-warp transpose lane c holds column c (bit r =
    // row r), so "window (j, m) is all zero in row r" is bit r of  ~(col_j | ... | col_{j+m}); the
    // neighbouring columns come from a per-warp shared-memory line and the OR is built incrementally.
    __shared__ uint32_t s_cw[CT_NWARP][96];
    int cnt[MAXM];      // lane b: rows (of the slow tiles) where window (32w+b, m) is all zero
#pragma unroll
    for (int m = 0; m < MAXM; ++m) cnt[m] = 0;
    int zero_rows = 0;  // rows of tiles in which every window of this word is trivially zero

    auto fetch = [&](int64_t r0, uint32_t& lo, uint32_t& mid, uint32_t& hi) {
        const int64_t r = r0 + lane;
        const bool live = r < r_end;
        lo = live ? c0[r] : 0xffffffffu;                   // dead rows count for nothing
        mid = (live && c1) ? c1[r] : 0u;
        hi = (live && c2) ? c2[r] : 0u;
    };
    // Sparse tiles (posterior samples: a handful of non-zero columns among the 64 a word's windows can
    // reach) skip the transposes: each non-zero column c is gathered with one ballot, in increasing
    // order, and lane j adds the rows it newly covers to E[c - j]; with E[m] = #rows first covered at
    // width m, the zero count of window (j, m) is 32 T - sum_{k <= m} E[k] over the T sparse tiles.
    constexpr bool SPARSE = MAXM <= 32;
    constexpr int SPARSE_MAX_COLS = 10;
    __shared__ int s_E[SPARSE ? CT_NWARP : 1][SPARSE ? 32 : 1][32];
    if (SPARSE) {
        for (int m = 0; m < 32; ++m) s_E[warp][m][lane] = 0;
        __syncwarp();
    }
    int sparse_tiles = 0;
    auto process = [&](const uint32_t lo, const uint32_t mid, const uint32_t hi) {
        if (!__any_sync(0xffffffffu, (lo | mid | hi) != 0u)) { zero_rows += 32; return; }
        if (SPARSE) {
            // only the first max_size - 1 columns of the next word are within reach of this word's windows
            const uint32_t reach = max_size >= 33 ? 0xffffffffu : ((1u << (max_size - 1)) - 1u);
            uint32_t nzl = __reduce_or_sync(0xffffffffu, lo), nzm = __reduce_or_sync(0xffffffffu, mid) & reach;
            if (__popc(nzl) + __popc(nzm) <= SPARSE_MAX_COLS) {
                uint32_t acc = 0u;                          // rows covered by the non-zero columns in [j, c)
                while (nzl) {                               // warp-uniform
                    const int c = __ffs(nzl) - 1;
                    nzl &= nzl - 1;
                    const uint32_t colc = __ballot_sync(0xffffffffu, (lo >> c) & 1u);
                    const int m = c - lane;
                    if (m >= 0 && m < max_size) {
                        s_E[warp][m][lane] += __popc(colc & ~acc);
                        acc |= colc;
                    }
                }
                while (nzm) {
                    const int c = __ffs(nzm) - 1;
                    nzm &= nzm - 1;
                    const uint32_t colc = __ballot_sync(0xffffffffu, (mid >> c) & 1u);
                    const int m = 32 + c - lane;            // >= 1
                    if (m < max_size) {
                        s_E[warp][m][lane] += __popc(colc & ~acc);
                        acc |= colc;
                    }
                }
                ++sparse_tiles;
                return;
            }
        }
        __syncwarp();
        uint32_t acc = warp_transpose32(lo);                // OR of columns j..j+m, one bit per row
        s_cw[warp][32 + lane] = warp_transpose32(mid);
        if (MAXM > 32) s_cw[warp][64 + lane] = warp_transpose32(hi);
        s_cw[warp][lane] = acc;
        __syncwarp();
        cnt[0] += __popc(~acc);
#pragma unroll
        for (int m = 1; m < MAXM; ++m) {
            if (m < max_size) {
                acc |= s_cw[warp][lane + m];
                cnt[m] += __popc(~acc);
            }
        }
    };
    // CS_PF tiles are requested ahead of the one being processed: the kernel is bound by the bytes a
    // warp keeps in flight, not by its instructions
    constexpr int CS_PF = 4;
    const int64_t step = 32 * CT_NWARP;
    uint32_t lo_n[CS_PF], mid_n[CS_PF], hi_n[CS_PF];
#pragma unroll
    for (int i = 0; i < CS_PF; ++i) fetch(r_begin + 32 * warp + i * step, lo_n[i], mid_n[i], hi_n[i]);
    for (int64_t rb = r_begin + 32 * warp; rb < r_end; rb += CS_PF * step) {
        uint32_t lo_c[CS_PF], mid_c[CS_PF], hi_c[CS_PF];
#pragma unroll
        for (int i = 0; i < CS_PF; ++i) { lo_c[i] = lo_n[i]; mid_c[i] = mid_n[i]; hi_c[i] = hi_n[i]; }
        if (rb + CS_PF * step < r_end) {
#pragma unroll
            for (int i = 0; i < CS_PF; ++i) fetch(rb + (CS_PF + i) * step, lo_n[i], mid_n[i], hi_n[i]);
        }
#pragma unroll
        for (int i = 0; i < CS_PF; ++i)
            if (rb + i * step < r_end) process(lo_c[i], mid_c[i], hi_c[i]);      // warp-uniform
    }
    if (SPARSE) {
        int covered = 0;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < max_size) {
                covered += s_E[warp][m][lane];
                cnt[m] += 32 * sparse_tiles - covered;
            }
        }
    }
    // reduce over the warps of the block, then one atomic per (m, column)
    const int64_t j = 32LL * w + lane;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        if (m < max_size) {                 // uniform
            __syncthreads();
            s_cnt[warp][lane] = cnt[m] + zero_rows;
            __syncthreads();
            if (warp == 0) {
                int t = 0;
#pragma unroll
                for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i][lane];
                if (j + m < p && t) atomicAdd(zero_counts + (int64_t)m * p + j, (unsigned long long)t);
            }
        }
    }
}

// ---- sequential zero counts, gap-histogram formulation ---------------------------
// Per row, a set bit at column c with g zero columns immediately before it is the FIRST non-zero
// column of exactly the windows starting at j = c - k, k = 0 .. g.  With
//     H[c][g] = #rows with bit c set and a preceding zero run of length g   (g capped at max_size - 1)
//     S[c][k] = sum_{g >= k} H[c][g]   = #rows whose first non-zero column at or after c - k is c
// the number of rows in which columns j .. j+m are all zero is
//     zero_counts[m][j] = N - sum_{k = 0 .. m} S[j + k][k].
// The scan does work only where a word is non-zero -- one warp-aggregated shared-memory increment
// per set bit -- instead of evaluating every (column, width) pair of every 32 x 32 tile, and any
// max_size is allowed (create_groups.py:321-332 accepts any).
template <bool SMEM_HIST>
__global__ void __launch_bounds__(CT_BLOCK) gap_hist_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int max_size, int64_t rows_per_block,
    unsigned long long* __restrict__ H /* [32 * words][max_size] */)
{
    extern __shared__ int s_h[];                        // [32][max_size]
    const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(N, r_begin + rows_per_block);
    const uint32_t* c0 = bits + (int64_t)w * N_pad;
    const int nh = 32 * max_size;
    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < nh; i += CT_BLOCK) s_h[i] = 0;
        __syncthreads();
    }
    unsigned long long* Hw = H + (int64_t)w * nh;
    constexpr int PF = 8;                               // row tiles a warp keeps in flight
    const int64_t step = 32 * CT_NWARP;
    const uint32_t* cp = (w > 0 && max_size > 1) ? bits + (int64_t)(w - 1) * N_pad : nullptr;
    // a lane that finds its word non-zero will need the zero run before it: the neighbouring word is
    // requested together with the prefetch, PF tiles ahead of its use
    uint32_t xn[PF], pn[PF];
    auto fetch = [&](int64_t r, uint32_t& x, uint32_t& pv) {
        x = r < r_end ? c0[r] : 0u;
        pv = (x != 0u && cp) ? cp[r] : 0u;
    };
#pragma unroll
    for (int i = 0; i < PF; ++i) fetch(r_begin + 32 * warp + i * step + lane, xn[i], pn[i]);
    for (int64_t rb = r_begin + 32 * warp; rb < r_end; rb += PF * step) {
        uint32_t xc[PF], pc[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) { xc[i] = xn[i]; pc[i] = pn[i]; }
#pragma unroll
        for (int i = 0; i < PF; ++i) fetch(rb + (PF + i) * step + lane, xn[i], pn[i]);
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            uint32_t x = xc[i];
            if (!__any_sync(0xffffffffu, x != 0u)) continue;            // the common tile: nothing set
            const int64_t r = rb + i * step + lane;
            // zero run that ends right before column 32 w of this row (only what the cap can see)
            int tz = 0;
            bool done = x == 0u || w == 0 || max_size == 1;
            if (!done) {
                if (pc[i]) { tz = __clz(pc[i]); done = true; } else tz = 32;
            }
            for (int k = 2; k <= w && 32 * (k - 1) < max_size - 1; ++k) {   // warp-uniform bounds; rare
                if (__all_sync(0xffffffffu, done)) break;
                if (!done) {
                    const uint32_t xp = bits[(int64_t)(w - k) * N_pad + r];
                    if (xp) { tz += __clz(xp); done = true; } else tz += 32;
                }
            }
            // Rows of a tile are consecutive kept sweeps of one chain and mostly carry the SAME word (the persistently
            // active columns): lanes are grouped ONCE by (word, zero run before it -- only what the cap can see), and
            // one lane per group books the group's set bits with the group's size as the weight.
            tz = min(tz, max_size - 1);
            const unsigned peers = __match_any_sync(0xffffffffu, ((unsigned long long)(unsigned)tz << 32) | x);
            if (x != 0u && lane == __ffs(peers) - 1) {
                const int weight = __popc(peers);
                int prev = -1 - tz;                     // the previous set bit, relative to column 32 w
                do {
                    const int b = __ffs(x) - 1;
                    x &= x - 1;
                    const int key = b * max_size + min(b - prev - 1, max_size - 1);
                    prev = b;
                    if (SMEM_HIST) atomicAdd(&s_h[key], weight);
                    else atomicAdd(Hw + key, (unsigned long long)weight);
                } while (x);
            }
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < nh; i += CT_BLOCK) {
            const int v = s_h[i];
            if (v) atomicAdd(Hw + i, (unsigned long long)v);
        }
    }
}

// H[c][g] -> S[c][k] = sum_{g >= k} H[c][g], in place; one thread per column
__global__ void gap_suffix_kernel(unsigned long long* __restrict__ H, int64_t ncols, int max_size)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    unsigned long long* h = H + c * max_size;
    unsigned long long acc = 0;
    for (int g = max_size - 1; g >= 0; --g) { acc += h[g]; h[g] = acc; }
}

// zero_counts[m][j] = N - sum_{k <= m} S[j + k][k]   (0 where the window leaves the matrix)
__global__ void gap_to_zero_counts_kernel(const unsigned long long* __restrict__ S, int64_t N, int64_t p,
                                          int max_size, unsigned long long* __restrict__ zero_counts)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    unsigned long long covered = 0;
    for (int m = 0; m < max_size; ++m) {
        if (j + m >= p) break;                      // the output was zero-filled
        covered += S[(j + m) * max_size + m];
        zero_counts[(int64_t)m * p + j] = (unsigned long long)N - covered;
    }
}

// ---- column-major copy of the bit matrix -----------------------------------
// One warp transposes 8 consecutive 32 x 32 tiles of word w (256 rows) and lane c writes the 8 row
// words of column 32 w + c as two 16-byte stores (one full 32-byte sector).
__global__ void __launch_bounds__(CT_BLOCK) transpose_bits_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int64_t NW, uint32_t* __restrict__ colbits)
{
    const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t t0 = 8 * ((int64_t)blockIdx.y * CT_NWARP + warp);           // first row word of this warp
    if (t0 >= NW) return;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = 32 * (t0 + i) + lane;
        v[i] = r < N ? bits[(int64_t)w * N_pad + r] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = warp_transpose32(v[i]);
    uint4* dst = reinterpret_cast<uint4*>(colbits + ((int64_t)32 * w + lane) * NW + t0);
    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// any-counts from the column-major copy: a group costs |group| N / 8 bytes instead of 4 N |group|.
// grid (groups, chunks); a thread ORs the members' row words four at a time.
__global__ void __launch_bounds__(CT_BLOCK) count_groups_colmajor_kernel(
    const uint32_t* __restrict__ colbits, int64_t NW, const int64_t* __restrict__ offsets,
    const int32_t* __restrict__ members, int64_t quads_per_block, unsigned long long* __restrict__ any_counts)
{
    __shared__ int s_cnt[CT_NWARP];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t m0 = offsets[g], m1 = offsets[g + 1];
    const int64_t nquad = NW / 4;
    const int64_t q_begin = (int64_t)blockIdx.y * quads_per_block, q_end = min(nquad, q_begin + quads_per_block);
    int cnt = 0;
    for (int64_t qd = q_begin + threadIdx.x; qd < q_end; qd += CT_BLOCK) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t k = m0; k < m1; ++k) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(colbits + (int64_t)members[k] * NW) + qd);
            acc.x |= x.x; acc.y |= x.y; acc.z |= x.z; acc.w |= x.w;
        }
        cnt += __popc(acc.x) + __popc(acc.y) + __popc(acc.z) + __popc(acc.w);
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i];
        if (t) atomicAdd(any_counts + g, (unsigned long long)t);
    }
}

// ---- arbitrary groups: any(samples[:, group]) ---------------------------
// grid (groups, chunks); lane = row.
__global__ void __launch_bounds__(CT_BLOCK) count_groups_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, const int64_t* __restrict__ offsets,
    const int32_t* __restrict__ members, int64_t rows_per_block, unsigned long long* __restrict__ any_counts)
{
    __shared__ int s_cnt[CT_NWARP];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t m0 = offsets[g], m1 = offsets[g + 1];
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(N, r_begin + rows_per_block);
    int cnt = 0;
    for (int64_t r = r_begin + threadIdx.x; r < r_end; r += CT_BLOCK) {
        uint32_t any = 0u;
        for (int64_t k = m0; k < m1; ++k) {
            const int cidx = members[k];
            any |= (bits[(int64_t)(cidx >> 5) * N_pad + r] >> (cidx & 31)) & 1u;
        }
        cnt += (int)any;
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i];
        if (t) atomicAdd(any_counts + g, (unsigned long long)t);
    }
}

// ---- pair (co-occurrence) counts -----------------------------------------
// pair_counts[a][b] = #rows with columns a and b both non-zero: the integer content of the
// reference's np.corrcoef(samples) (create_groups.py:523-535).  grid (word pairs wa <= wb, chunks);
// a warp transposes a 32-row tile of both words (lane = column) and walks the 32 cyclic diagonals,
// so lane a accumulates the pairs (a, a+s mod 32) without any broadcast.
__global__ void __launch_bounds__(CT_BLOCK) count_pairs_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, int64_t p, int words,
    int64_t rows_per_block, unsigned long long* __restrict__ pair_counts)
{
    __shared__ int s_cnt[32][33];
    // unrank the word pair (wa <= wb) from blockIdx.x
    int wa = 0, rem = blockIdx.x;
    while (rem >= words - wa) { rem -= words - wa; ++wa; }
    const int wb = wa + rem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(N, r_begin + rows_per_block);
    const uint32_t* ca = bits + (int64_t)wa * N_pad;
    const uint32_t* cb = bits + (int64_t)wb * N_pad;
    int cnt[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) cnt[s] = 0;
    for (int64_t r0 = r_begin + 32 * warp; r0 < r_end; r0 += 32 * CT_NWARP) {
        const int64_t r = r0 + lane;
        const uint32_t xa = r < r_end ? ca[r] : 0u;
        const uint32_t xb = r < r_end ? cb[r] : 0u;
        if (!__any_sync(0xffffffffu, xa != 0u) || !__any_sync(0xffffffffu, xb != 0u)) continue;
        const uint32_t ta = warp_transpose32(xa);
        const uint32_t tb = wa == wb ? ta : warp_transpose32(xb);
#pragma unroll
        for (int s = 0; s < 32; ++s) cnt[s] += __popc(ta & __shfl_sync(0xffffffffu, tb, (lane + s) & 31));
    }
    // block reduction through shared memory: s_cnt[a][b]
    for (int e = threadIdx.x; e < 32 * 33; e += CT_BLOCK) (&s_cnt[0][0])[e] = 0;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 32; ++s) if (cnt[s]) atomicAdd(&s_cnt[lane][(lane + s) & 31], cnt[s]);
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += CT_BLOCK) {
        const int a = e >> 5, b = e & 31;
        const int64_t ja = 32LL * wa + a, jb = 32LL * wb + b;
        const int v = s_cnt[a][b];
        if (v && ja < p && jb < p) {
            atomicAdd(pair_counts + ja * p + jb, (unsigned long long)v);
            if (wa != wb) atomicAdd(pair_counts + jb * p + ja, (unsigned long long)v);
        }
    }
}

// ---- exact FWER of a selection: rows in which SOME selected group holds no signal ------------
// (blip.py:270-275: false_disc |= all(samples[:, group] == 0, axis=1), then the mean).
// grid (chunks); lane = row; every thread walks all selected groups for its row and stops at the
// first empty one.  Group lists are small (a selection), member words are re-read from L1/L2.
__global__ void __launch_bounds__(CT_BLOCK) count_false_rows_kernel(
    const uint32_t* __restrict__ bits, int64_t N, int64_t N_pad, const int64_t* __restrict__ offsets,
    const int32_t* __restrict__ members, int n_groups, unsigned long long* __restrict__ out)
{
    __shared__ int s_cnt[CT_NWARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cnt = 0;
    for (int64_t r = (int64_t)blockIdx.x * CT_BLOCK + threadIdx.x; r < N; r += (int64_t)gridDim.x * CT_BLOCK) {
        bool false_disc = false;
        for (int g = 0; g < n_groups && !false_disc; ++g) {
            uint32_t any = 0u;
            for (int64_t k = offsets[g]; k < offsets[g + 1] && !any; ++k) {
                const int cidx = members[k];
                any = (bits[(int64_t)(cidx >> 5) * N_pad + r] >> (cidx & 31)) & 1u;
            }
            false_disc = !any;
        }
        cnt += false_disc ? 1 : 0;
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i];
        if (t) atomicAdd(out, (unsigned long long)t);
    }
}

// The same scan on the column-major copy, 128 rows per thread step: a row is a false discovery as
// soon as one selected group has no member set in it.
__global__ void __launch_bounds__(CT_BLOCK) count_false_rows_colmajor_kernel(
    const uint32_t* __restrict__ colbits, int64_t N, int64_t NW, const int64_t* __restrict__ offsets,
    const int32_t* __restrict__ members, int n_groups, unsigned long long* __restrict__ out)
{
    __shared__ int s_cnt[CT_NWARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nquad = NW / 4;
    int cnt = 0;
    for (int64_t qd = (int64_t)blockIdx.x * CT_BLOCK + threadIdx.x; qd < nquad; qd += (int64_t)gridDim.x * CT_BLOCK) {
        uint4 fd = make_uint4(0u, 0u, 0u, 0u);
        for (int g = 0; g < n_groups; ++g) {
            uint4 any = make_uint4(0u, 0u, 0u, 0u);
            for (int64_t k = offsets[g]; k < offsets[g + 1]; ++k) {
                const uint4 x = __ldg(reinterpret_cast<const uint4*>(colbits + (int64_t)members[k] * NW) + qd);
                any.x |= x.x; any.y |= x.y; any.z |= x.z; any.w |= x.w;
            }
            fd.x |= ~any.x; fd.y |= ~any.y; fd.z |= ~any.z; fd.w |= ~any.w;
            if ((fd.x & fd.y & fd.z & fd.w) == 0xffffffffu) break;
        }
        // rows >= N do not exist
        const int64_t r0 = 128 * qd;
        uint32_t w4[4] = { fd.x, fd.y, fd.z, fd.w };
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t left = N - (r0 + 32 * i);
            if (left <= 0) w4[i] = 0u; else if (left < 32) w4[i] &= (1u << left) - 1u;
            cnt += __popc(w4[i]);
        }
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < CT_NWARP; ++i) t += s_cnt[i];
        if (t) atomicAdd(out, (unsigned long long)t);
    }
}

int alloc_bits(blipgpu_ctx* ctx, int64_t N, int64_t p, blipgpu_bits** out)
{
    blipgpu_bits* b = new (std::nothrow) blipgpu_bits();
    if (!b) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    b->ctx = ctx; b->N = N; b->N_pad = round_up(std::max<int64_t>(N, 1), 32); b->p = p;
    b->words = (p + 31) / 32; b->owns = true; b->bits = nullptr; b->colbits = nullptr; b->NW = 0;
    const size_t bytes = sizeof(uint32_t) * std::max<int64_t>(b->words, 1) * b->N_pad;
    cudaError_t e = dev_alloc(ctx, (void**)&b->bits, bytes);
    if (e != cudaSuccess) { delete b; blip_set_error("bits allocation (%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    e = cudaMemsetAsync(b->bits, 0, bytes, ctx->stream);
    if (e != cudaSuccess) { dev_free(ctx, b->bits); delete b; blip_set_error("memset failed: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    *out = b;
    return BLIP_OK;
}

int64_t pick_rows_per_block(int64_t N, int64_t columns_of_blocks, int sm_count)
{
    // enough blocks to fill the machine a few times, but >= 2048 rows each so that the
    // per-block atomics stay negligible
    const int64_t target_blocks = 8LL * sm_count;
    int64_t chunks = std::max<int64_t>(1, target_blocks / std::max<int64_t>(1, columns_of_blocks));
    int64_t rpb = round_up((N + chunks - 1) / chunks, 32 * CT_NWARP);
    rpb = std::max<int64_t>(rpb, 2048);
    return rpb;
}

// run `launch(dev_counts)` into either the caller's device buffer or a temporary
template <class F>
int with_output(blipgpu_ctx* ctx, int64_t count, int64_t* out, int out_is_device, F launch)
{
    unsigned long long* dev = nullptr;
    bool temp = !out_is_device;
    if (temp) {
        cudaError_t e = dev_alloc(ctx, (void**)&dev, sizeof(unsigned long long) * std::max<int64_t>(count, 1));
        if (e != cudaSuccess) { blip_set_error("count buffer allocation failed: %s", cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    } else dev = reinterpret_cast<unsigned long long*>(out);
    int rc = BLIP_OK;
    cudaError_t e = cudaMemsetAsync(dev, 0, sizeof(unsigned long long) * count, ctx->stream);
    if (e == cudaSuccess) {
        ScopedTimer timer(ctx, false);
        launch(dev);
        e = cudaGetLastError();
        timer.stop(1);
    }
    if (e == cudaSuccess && temp) e = cudaMemcpyAsync(out, dev, sizeof(int64_t) * count, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { blip_set_error("counting kernel failed: %s", cudaGetErrorString(e)); rc = BLIP_ERR_CUDA; }
    if (temp) dev_free(ctx, dev);
    return rc;
}

} // namespace

int blip_bits_alloc(blipgpu_ctx* ctx, int64_t N, int64_t p, blipgpu_bits** out) { return alloc_bits(ctx, N, p, out); }

extern "C" int blipgpu_bits_from_samples(blipgpu_samples* s, int32_t zero_first_column, blipgpu_bits** out)
{
    if (!s || !out) { blip_set_error("bits_from_samples: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    blipgpu_bits* b = nullptr;
    BLIP_TRY(alloc_bits(s->ctx, s->rows, s->p, &b));
    cudaStream_t st = s->ctx->stream;
    // same layout and the same row padding (rows_pad == N_pad): one device copy
    cudaError_t e = cudaMemcpyAsync(b->bits, s->bits, sizeof(uint32_t) * s->words * s->rows_pad, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && zero_first_column)
        clear_first_column_kernel<<<(unsigned)((s->rows + 255) / 256), 256, 0, st>>>(b->bits, s->rows);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { blipgpu_bits_free(b); blip_set_error("bits_from_samples: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    *out = b;
    return BLIP_OK;
}

extern "C" int blipgpu_bits_clear_first_column(blipgpu_bits* b)
{
    if (!b) { blip_set_error("bits_clear_first_column: NULL handle"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    if (b->N == 0) return BLIP_OK;
    cudaStream_t st = b->ctx->stream;
    clear_first_column_kernel<<<(unsigned)((b->N + 255) / 256), 256, 0, st>>>(b->bits, b->N);
    CUDA_TRY(cudaGetLastError());
    if (b->colbits) {                       // the column-major copy holds column 0 in its first NW words
        CUDA_TRY(cudaMemsetAsync(b->colbits, 0, sizeof(uint32_t) * (size_t)b->NW, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    b->ctx->kernel_launches += 1;
    return BLIP_OK;
}

extern "C" int blipgpu_bits_from_dense(blipgpu_ctx* ctx, const void* samples, int32_t dtype, int64_t N,
                                       int64_t p, blipgpu_bits** out)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!out || N < 0 || p <= 0 || (N > 0 && !samples) || p >= (int64_t)1 << 31) { blip_set_error("bits_from_dense: bad arguments"); return BLIP_ERR_ARG; }
    if (dtype != BLIPGPU_DTYPE_F64 && dtype != BLIPGPU_DTYPE_U8) { blip_set_error("bits_from_dense: unknown dtype %d", dtype); return BLIP_ERR_ARG; }
    blipgpu_bits* b = nullptr;
    BLIP_TRY(alloc_bits(ctx, N, p, &b));
    if (N == 0) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); *out = b; return BLIP_OK; }
    const size_t esz = dtype == BLIPGPU_DTYPE_F64 ? 8 : 1;
    const int64_t stage_rows = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)(256 << 20) / (int64_t)(esz * p)));
    void* stage = nullptr;
    cudaError_t e = dev_alloc(ctx, &stage, esz * stage_rows * p);
    if (e != cudaSuccess) { blipgpu_bits_free(b); blip_set_error("bits_from_dense: staging allocation failed: %s", cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    cudaStream_t st = ctx->stream;
    for (int64_t r = 0; r < N && e == cudaSuccess; r += stage_rows) {
        const int64_t nr = std::min<int64_t>(stage_rows, N - r);
        e = cudaMemcpyAsync(stage, (const char*)samples + esz * r * p, esz * nr * p, cudaMemcpyHostToDevice, st);
        const int64_t nwarps = nr * b->words;
        const unsigned grid = (unsigned)((nwarps + 7) / 8);
        if (dtype == BLIPGPU_DTYPE_F64)
            pack_dense_kernel<double><<<grid, 256, 0, st>>>((const double*)stage, nr, p, (int)b->words, r, b->N_pad, b->bits);
        else
            pack_dense_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)stage, nr, p, (int)b->words, r, b->N_pad, b->bits);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(ctx, stage);
    if (e != cudaSuccess) { blipgpu_bits_free(b); blip_set_error("bits_from_dense: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    *out = b;
    return BLIP_OK;
}

extern "C" int blipgpu_bits_shape(blipgpu_bits* b, int64_t* N, int64_t* p)
{
    if (!b) { blip_set_error("bits_shape: NULL handle"); return BLIP_ERR_ARG; }
    if (N) *N = b->N;
    if (p) *p = b->p;
    return BLIP_OK;
}

extern "C" int blipgpu_bits_select_columns(blipgpu_bits* b, const int64_t* cols, int64_t ncols, blipgpu_bits** out)
{
    if (!b || !out || ncols <= 0 || !cols) { blip_set_error("select_columns: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    std::vector<int32_t> c32((size_t)ncols);
    for (int64_t i = 0; i < ncols; ++i) {
        if (cols[i] < 0 || cols[i] >= b->p) { blip_set_error("select_columns: column %lld out of range", (long long)cols[i]); return BLIP_ERR_ARG; }
        c32[(size_t)i] = (int32_t)cols[i];
    }
    blipgpu_bits* o = nullptr;
    BLIP_TRY(alloc_bits(b->ctx, b->N, ncols, &o));
    int32_t* d_cols = nullptr;
    cudaStream_t st = b->ctx->stream;
    cudaError_t e = dev_alloc(b->ctx, (void**)&d_cols, sizeof(int32_t) * ncols);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cols, c32.data(), sizeof(int32_t) * ncols, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && b->N > 0) {
        dim3 grid((unsigned)((b->N + 255) / 256), (unsigned)o->words);
        select_columns_kernel<<<grid, 256, 0, st>>>(b->bits, b->N, b->N_pad, d_cols, (int)ncols, (int)o->words, o->bits);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(b->ctx, d_cols);
    if (e != cudaSuccess) { blipgpu_bits_free(o); blip_set_error("select_columns: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    *out = o;
    return BLIP_OK;
}

extern "C" int blipgpu_bits_to_dense(blipgpu_bits* b, uint8_t* out)
{
    if (!b || (!out && b->N > 0)) { blip_set_error("bits_to_dense: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    if (b->N == 0) return BLIP_OK;
    const int64_t stage_rows = std::min<int64_t>(b->N, std::max<int64_t>(1, (int64_t)(256 << 20) / b->p));
    uint8_t* stage = nullptr;
    CUDA_TRY(dev_alloc(b->ctx, (void**)&stage, (size_t)stage_rows * b->p));
    cudaStream_t st = b->ctx->stream;
    cudaError_t e = cudaSuccess;
    for (int64_t r = 0; r < b->N && e == cudaSuccess; r += stage_rows) {
        const int64_t nr = std::min<int64_t>(stage_rows, b->N - r);
        unpack_kernel<<<(unsigned)((nr * b->p + 255) / 256), 256, 0, st>>>(b->bits, b->N, b->N_pad, b->p, r, nr, stage);
        e = cudaMemcpyAsync(out + r * b->p, stage, (size_t)nr * b->p, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(b->ctx, stage);
    if (e != cudaSuccess) { blip_set_error("bits_to_dense: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    return BLIP_OK;
}

extern "C" int blipgpu_count_columns(blipgpu_bits* b, int64_t* counts, int32_t out_is_device)
{
    if (!b || !counts) { blip_set_error("count_columns: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    const int64_t rpb = pick_rows_per_block(b->N, b->words, b->ctx->sm_count);
    const unsigned chunks = (unsigned)std::max<int64_t>(1, (b->N + rpb - 1) / rpb);
    return with_output(b->ctx, b->p, counts, out_is_device, [&](unsigned long long* dev) {
        if (b->N > 0)
            count_columns_kernel<<<dim3((unsigned)b->words, chunks), CT_BLOCK, 0, b->ctx->stream>>>(b->bits, b->N, b->N_pad, b->p, rpb, dev);
    });
}

extern "C" int blipgpu_count_sequential(blipgpu_bits* b, int32_t max_size, int64_t* zero_counts, int32_t out_is_device)
{
    if (!b || !zero_counts || max_size <= 0) { blip_set_error("count_sequential: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    // windows wider than the matrix do not exist: their rows of the output stay zero
    const int eff = (int)std::min<int64_t>(max_size, b->p);
    const int64_t rpb = pick_rows_per_block(b->N, b->words, b->ctx->sm_count);
    const unsigned chunks = (unsigned)std::max<int64_t>(1, (b->N + rpb - 1) / rpb);
    const bool legacy = getenv("BLIPGPU_COUNT_SEQ_TILES") != nullptr && max_size <= 64;   // the round-1 tile kernel, for A/B timing
    // The work of the gap histogram sits where the set bits are -- for posterior samples in the few words
    // that hold the persistently active columns -- so the rows are cut into many short chunks: the CTAs of
    // a heavy word then spread over all SMs instead of queueing on a few (0.78 ms with three chunks).
    const int64_t rpb_gap = 16384;      // 2048 / 4096 / 8192 / 16384 rows measured: 0.54 / 0.44 / 0.39 / 0.38 ms at C2
    const unsigned chunks_gap = (unsigned)std::max<int64_t>(1, (b->N + rpb_gap - 1) / rpb_gap);
    DevPoolScope pool_scope(b->ctx->stream);
    DevBuf hist;
    if (!legacy && b->N > 0) BLIP_TRY(hist.alloc(sizeof(unsigned long long) * 32 * (size_t)b->words * (size_t)eff));
    return with_output(b->ctx, (int64_t)max_size * b->p, zero_counts, out_is_device, [&](unsigned long long* dev) {
        if (b->N == 0) return;
        dim3 grid((unsigned)b->words, chunks);
        cudaStream_t st = b->ctx->stream;
        if (legacy) {
            if (max_size <= 32)
                count_sequential_kernel<32><<<grid, CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, b->p, (int)b->words, max_size, rpb, dev);
            else
                count_sequential_kernel<64><<<grid, CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, b->p, (int)b->words, max_size, rpb, dev);
            return;
        }
        unsigned long long* H = hist.as<unsigned long long>();
        cudaMemsetAsync(H, 0, hist.bytes, st);
        const size_t smem = sizeof(int) * 32 * (size_t)eff;
        dim3 ggrid((unsigned)b->words, chunks_gap);
        if (smem <= 40 * 1024) gap_hist_kernel<true><<<ggrid, CT_BLOCK, smem, st>>>(b->bits, b->N, b->N_pad, eff, rpb_gap, H);
        else gap_hist_kernel<false><<<ggrid, CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, eff, rpb_gap, H);
        const int64_t ncols = 32 * b->words;
        gap_suffix_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(H, ncols, eff);
        gap_to_zero_counts_kernel<<<(unsigned)((b->p + 127) / 128), 128, 0, st>>>(H, b->N, b->p, eff, dev);
        b->ctx->kernel_launches += 2;
    });
}

static void ensure_colbits(blipgpu_bits* b, cudaStream_t st);

// Tensor-core path (pairs_tc.cuh): u8 x u8 -> s32 GEMM over the column-major bit matrix, any number of columns.
// Chosen from 512 columns up (below that the scalar kernel's work is a few dozen word pairs); BLIPGPU_PAIRS_TC=1
// forces it, BLIPGPU_PAIRS_TC=0 forbids it.
static int count_pairs_tc(blipgpu_bits* b, int64_t* pair_counts, int32_t out_is_device)
{
    cudaStream_t st = b->ctx->stream;
    const int64_t C = 32 * b->words;
    const int n_mt = (int)((C + PT_M - 1) / PT_M), n_nt = (int)((C + PT_N - 1) / PT_N);
    int64_t tiles = 0;
    for (int mt = 0; mt < n_mt; ++mt) tiles += n_nt - (mt * PT_M) / PT_N;
    const int64_t n_octs = b->NW / 8;
    int64_t ksplit = std::max<int64_t>(1, std::min<int64_t>(n_octs, (2LL * b->ctx->sm_count + tiles - 1) / tiles));
    const int64_t octs_per_split = (n_octs + ksplit - 1) / ksplit;
    ksplit = (n_octs + octs_per_split - 1) / octs_per_split;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(count_pairs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM) != cudaSuccess) {
            blip_set_error("count_pairs: cannot reserve %d bytes of shared memory", PT_SMEM); return BLIP_ERR_CUDA;
        }
        attr_set = true;
    }
    return with_output(b->ctx, b->p * b->p, pair_counts, out_is_device, [&](unsigned long long* dev) {
        count_pairs_tc_kernel<<<dim3((unsigned)tiles, (unsigned)ksplit), PT_THREADS, PT_SMEM, st>>>(
            b->colbits, b->NW, C, b->p, n_nt, octs_per_split, n_octs, dev);
    });
}

extern "C" int blipgpu_count_pairs(blipgpu_bits* b, int64_t* pair_counts, int32_t out_is_device)
{
    if (!b || !pair_counts) { blip_set_error("count_pairs: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    {
        const char* env = getenv("BLIPGPU_PAIRS_TC");
        const bool want_tc = env ? env[0] == '1' : b->p >= 512;
        if (want_tc && b->N > 0 && b->p > 0) {
            ensure_colbits(b, b->ctx->stream);
            if (b->colbits) return count_pairs_tc(b, pair_counts, out_is_device);
        }
    }
    if (b->p > 16384) { blip_set_error("count_pairs: more than 16384 columns need the tensor-core path (column-major copy could not be allocated)"); return BLIP_ERR_UNSUPPORTED; }
    const int64_t npair = b->words * (b->words + 1) / 2;
    const int64_t rpb = pick_rows_per_block(b->N, npair, b->ctx->sm_count);
    const unsigned chunks = (unsigned)std::max<int64_t>(1, (b->N + rpb - 1) / rpb);
    return with_output(b->ctx, b->p * b->p, pair_counts, out_is_device, [&](unsigned long long* dev) {
        if (b->N == 0 || b->p == 0) return;
        count_pairs_kernel<<<dim3((unsigned)npair, chunks), CT_BLOCK, 0, b->ctx->stream>>>(
            b->bits, b->N, b->N_pad, b->p, (int)b->words, rpb, dev);
    });
}

// Builds the column-major copy of the bit matrix once per handle (no-op if it exists or if there is
// no room for it: the callers then stay on their row-major kernels).
static void ensure_colbits(blipgpu_bits* b, cudaStream_t st)
{
    if (b->colbits || b->N <= 0) return;
    const int64_t NW = round_up((b->N + 31) / 32, 8);
    uint32_t* cb = nullptr;
    if (dev_alloc(b->ctx, (void**)&cb, sizeof(uint32_t) * 32 * b->words * NW) != cudaSuccess) { cudaGetLastError(); return; }
    const unsigned ty = (unsigned)((NW / 8 + CT_NWARP - 1) / CT_NWARP);
    transpose_bits_kernel<<<dim3((unsigned)b->words, ty), CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, NW, cb);
    b->ctx->kernel_launches += 1;
    b->colbits = cb; b->NW = NW;
}

extern "C" int blipgpu_count_groups(blipgpu_bits* b, const int64_t* offsets, const int64_t* members,
                                    int64_t n_groups, int64_t* any_counts, int32_t out_is_device)
{
    if (!b || n_groups < 0 || (n_groups > 0 && (!offsets || !members || !any_counts))) { blip_set_error("count_groups: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    if (n_groups == 0) return BLIP_OK;
    const int64_t total = offsets[n_groups];
    std::vector<int32_t> m32((size_t)std::max<int64_t>(total, 1));
    for (int64_t i = 0; i < total; ++i) {
        if (members[i] < 0 || members[i] >= b->p) { blip_set_error("count_groups: member %lld out of range", (long long)members[i]); return BLIP_ERR_ARG; }
        m32[(size_t)i] = (int32_t)members[i];
    }
    int64_t* d_off = nullptr; int32_t* d_mem = nullptr;
    cudaStream_t st = b->ctx->stream;
    cudaError_t e = dev_alloc(b->ctx, (void**)&d_off, sizeof(int64_t) * (n_groups + 1));
    if (e == cudaSuccess) e = dev_alloc(b->ctx, (void**)&d_mem, sizeof(int32_t) * std::max<int64_t>(total, 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && total > 0) e = cudaMemcpyAsync(d_mem, m32.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { dev_free(b->ctx, d_off); dev_free(b->ctx, d_mem); blip_set_error("count_groups: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    const int64_t rpb = pick_rows_per_block(b->N, n_groups, b->ctx->sm_count);
    const unsigned chunks = (unsigned)std::max<int64_t>(1, (b->N + rpb - 1) / rpb);
    // With enough members to pay for it (the transposition moves 2 N p / 8 bytes once per handle) the
    // counts come from a column-major copy of the bit matrix: |group| N / 8 bytes per group instead
    // of one 32-bit word per (row, member).
    if (total > b->p / 8) ensure_colbits(b, st);
    int rc = with_output(b->ctx, n_groups, any_counts, out_is_device, [&](unsigned long long* dev) {
        if (b->N == 0) return;
        // grid.x is limited to 2^31-1 groups; grid.y to 65535 chunks
        if (b->colbits) {
            const int64_t nquad = b->NW / 4;
            const int64_t want = std::max<int64_t>(1, 8LL * b->ctx->sm_count / std::max<int64_t>(1, n_groups));
            const int64_t qpb = std::max<int64_t>(CT_BLOCK, round_up((nquad + want - 1) / want, CT_BLOCK));
            const unsigned cy = (unsigned)std::max<int64_t>(1, (nquad + qpb - 1) / qpb);
            count_groups_colmajor_kernel<<<dim3((unsigned)n_groups, cy), CT_BLOCK, 0, st>>>(b->colbits, b->NW, d_off, d_mem, qpb, dev);
        } else {
            count_groups_kernel<<<dim3((unsigned)n_groups, chunks), CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, d_off, d_mem, rpb, dev);
        }
    });
    dev_free(b->ctx, d_off); dev_free(b->ctx, d_mem);
    return rc;
}

extern "C" int blipgpu_count_false_rows(blipgpu_bits* b, const int64_t* offsets, const int64_t* members,
                                        int64_t n_groups, int64_t* count, int32_t out_is_device)
{
    if (!b || n_groups < 0 || !count || (n_groups > 0 && (!offsets || !members))) { blip_set_error("count_false_rows: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(b->ctx));
    const int64_t total = n_groups > 0 ? offsets[n_groups] : 0;
    std::vector<int32_t> m32((size_t)std::max<int64_t>(total, 1));
    for (int64_t i = 0; i < total; ++i) {
        if (members[i] < 0 || members[i] >= b->p) { blip_set_error("count_false_rows: member %lld out of range", (long long)members[i]); return BLIP_ERR_ARG; }
        m32[(size_t)i] = (int32_t)members[i];
    }
    int64_t* d_off = nullptr; int32_t* d_mem = nullptr;
    cudaStream_t st = b->ctx->stream;
    cudaError_t e = dev_alloc(b->ctx, (void**)&d_off, sizeof(int64_t) * (n_groups + 1));
    if (e == cudaSuccess) e = dev_alloc(b->ctx, (void**)&d_mem, sizeof(int32_t) * std::max<int64_t>(total, 1));
    const int64_t zero = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, n_groups > 0 ? offsets : &zero, sizeof(int64_t) * (n_groups + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && total > 0) e = cudaMemcpyAsync(d_mem, m32.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { dev_free(b->ctx, d_off); dev_free(b->ctx, d_mem); blip_set_error("count_false_rows: %s", cudaGetErrorString(e)); return BLIP_ERR_CUDA; }
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((b->N + CT_BLOCK - 1) / CT_BLOCK, 8LL * b->ctx->sm_count));
    int rc = with_output(b->ctx, 1, count, out_is_device, [&](unsigned long long* dev) {
        if (b->N == 0 || n_groups == 0) return;
        if (b->colbits) {
            const unsigned gq = (unsigned)std::max<int64_t>(1, std::min<int64_t>((b->NW / 4 + CT_BLOCK - 1) / CT_BLOCK, 8LL * b->ctx->sm_count));
            count_false_rows_colmajor_kernel<<<gq, CT_BLOCK, 0, st>>>(b->colbits, b->N, b->NW, d_off, d_mem, (int)n_groups, dev);
        } else {
            count_false_rows_kernel<<<grid, CT_BLOCK, 0, st>>>(b->bits, b->N, b->N_pad, d_off, d_mem, (int)n_groups, dev);
        }
    });
    dev_free(b->ctx, d_off); dev_free(b->ctx, d_mem);
    return rc;
}

extern "C" int blipgpu_bits_free(blipgpu_bits* b)
{
    if (!b) return BLIP_OK;
    cudaSetDevice(b->ctx->device);
    if (b->owns) dev_free(b->ctx, b->bits);
    if (b->colbits) dev_free(b->ctx, b->colbits);
    delete b;
    return BLIP_OK;
}
