// pairs_tc.cuh -- co-occurrence counts of ALL column pairs on the 5th-generation tensor cores.
//
// pair_counts[a][b] = #rows in which columns a and b are both non-zero = (S^T S)[a][b] for the 0/1 indicator
// matrix S: the integer content of the reference's np.corrcoef(samples) (pyblip/create_groups.py:523-535).
// This is a contraction over the rows, so it runs as an integer GEMM: tcgen05.mma kind::i8 (u8 x u8 -> s32),
// both operands K-major in shared memory, the 128 x 256 accumulator tile in tensor memory.
//
// The operands never exist in global memory as bytes.  The sample store keeps a column-major copy of the bit
// matrix (colbits[c][t] = 32 rows of column c); every producer thread owns one operand row (= one column of S),
// reads 32 bytes of its bit column (256 rows), expands each 32-bit word to 32 bytes (0 / 1) with four
// multiply-and-mask steps per 16 bits, and stores them straight into the canonical K-major core-matrix layout
// (8 rows x 16 bytes per core matrix, no swizzle) that the UMMA shared-memory descriptor names.  A stage holds
// four k-steps (128 rows) of the A rows (128 columns of S) and the B rows (256 columns of S):
//
//   producers (12 warps)  wait empty[stage] -> expand -> fence.proxy.async -> arrive full[stage]
//   MMA issuer (1 thread) wait full[stage] -> 4 x tcgen05.mma (M 128, N 256, K 32) -> tcgen05.commit empty[stage]
//   epilogue (warps 0-3)  wait the last commit -> tcgen05.ld 32 lanes x 32 columns -> add into the int64 result
//                         (the upper triangle a <= b and its mirror)
//
// grid: (tiles that touch the upper triangle, k-splits).  Exact: products of 0/1 bytes accumulate in int32
// (rows per k-split < 2^31), the k-splits add up in int64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int PT_M = 128, PT_N = 256;
constexpr int PT_ROWS = PT_M + PT_N;             // operand rows (columns of S) per CTA = producer threads
constexpr int PT_KS = 4;                          // k-steps (32 rows of S each) per stage
constexpr int PT_STAGES = 3;
constexpr int PT_A_KSTEP = PT_M * 32, PT_B_KSTEP = PT_N * 32;      // bytes of one k-step of an operand
constexpr int PT_STAGE_BYTES = (PT_A_KSTEP + PT_B_KSTEP) * PT_KS;  // 48 KB
constexpr int PT_THREADS = PT_ROWS + 32;          // + the MMA warp
constexpr int PT_SMEM = PT_STAGES * PT_STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t pt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pt_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pt_mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(pt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pt_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(pt_smem_u32(bar)), "r"(parity) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); the two 16-byte K chunks of a
// k-step are 128 bytes apart (leading byte offset), groups of 8 rows 256 bytes apart (stride byte offset)
__device__ __forceinline__ uint64_t pt_smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)(128 >> 4) << 16;                  // leading byte offset, bits [16,30)
    d |= (uint64_t)(256 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    return d;                                         // base offset 0, layout type 0 = SWIZZLE_NONE
}
// kind::i8 instruction descriptor: D = s32, A = B = u8, both K-major, N = 256, M = 128
constexpr uint32_t PT_IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(PT_N >> 3) << 17) | ((uint32_t)(PT_M >> 4) << 24);

// 16 bits -> 16 bytes of 0 / 1
__device__ __forceinline__ uint4 pt_expand16(uint32_t h)
{
    uint4 r;
    r.x = ((h & 0xFu) * 0x00204081u) & 0x01010101u;
    r.y = (((h >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.z = (((h >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.w = (((h >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return r;
}

// colbits [32 * words][NW]; k-split y covers word quads [y * quads_per_split, ...) of every column
__global__ void __launch_bounds__(PT_THREADS, 1) count_pairs_tc_kernel(
    const uint32_t* __restrict__ colbits, int64_t NW, int64_t ncols_alloc, int64_t p, int n_nt,
    int64_t octs_per_split, int64_t n_octs, unsigned long long* __restrict__ pair_counts)
{
    extern __shared__ __align__(1024) uint8_t pt_sm[];
    __shared__ uint64_t full_bar[PT_STAGES], empty_bar[PT_STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // unrank the tile: tiles are listed row by row (mt), each row from its first N tile that reaches the diagonal
    int mt = 0, rem = blockIdx.x;
    for (;;) {
        const int first_nt = (mt * PT_M) / PT_N;
        const int cnt = n_nt - first_nt;
        if (rem < cnt) { rem += first_nt; break; }
        rem -= cnt; ++mt;
    }
    const int nt = rem;
    const int64_t oct0 = (int64_t)blockIdx.y * octs_per_split;
    const int64_t oct1 = min(n_octs, oct0 + octs_per_split);
    const int n_stage = (int)(2 * (oct1 - oct0));                  // an oct = 8 words = 2 stages

    uint8_t* stage_base = pt_sm;
    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) { pt_mbar_init(&full_bar[s], PT_ROWS); pt_mbar_init(&empty_bar[s], 1); }
        pt_mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PT_ROWS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pt_smem_u32(&tmem_base_s)), "r"(PT_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (tid < PT_ROWS) {
        // ===== producers: one operand row (one column of S) each =====
        const bool is_a = tid < PT_M;
        const int r = is_a ? tid : tid - PT_M;
        const int64_t col = is_a ? (int64_t)mt * PT_M + r : (int64_t)nt * PT_N + r;
        const bool valid = col < ncols_alloc;
        const uint32_t* src = colbits + (valid ? col : 0) * NW;
        const uint32_t op_off = (is_a ? 0u : (uint32_t)(PT_A_KSTEP * PT_KS)) + (uint32_t)((r >> 3) * 256 + (r & 7) * 16);
        const uint32_t ks_stride = is_a ? (uint32_t)PT_A_KSTEP : (uint32_t)PT_B_KSTEP;
        int it = 0;
        // the 32 bytes of the next oct are requested before the current ones are expanded (an L2 / HBM round trip
        // per two stages would otherwise sit in front of every expansion)
        uint4 nx0 = make_uint4(0u, 0u, 0u, 0u), nx1 = nx0;
        if (valid && oct0 < oct1) {
            nx0 = __ldg(reinterpret_cast<const uint4*>(src + oct0 * 8));
            nx1 = __ldg(reinterpret_cast<const uint4*>(src + oct0 * 8 + 4));
        }
        for (int64_t oct = oct0; oct < oct1; ++oct) {
            const uint4 w0 = nx0, w1 = nx1;
            if (valid && oct + 1 < oct1) {
                nx0 = __ldg(reinterpret_cast<const uint4*>(src + (oct + 1) * 8));
                nx1 = __ldg(reinterpret_cast<const uint4*>(src + (oct + 1) * 8 + 4));
            }
#pragma unroll
            for (int half = 0; half < 2; ++half, ++it) {
                const int slot = it % PT_STAGES;
                const uint32_t phase = (uint32_t)((it / PT_STAGES) & 1);
                pt_mbar_wait(&empty_bar[slot], phase ^ 1u);
                uint8_t* dst = stage_base + (size_t)slot * PT_STAGE_BYTES + op_off;
                const uint4 w = half == 0 ? w0 : w1;
                const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                for (int ks = 0; ks < PT_KS; ++ks) {
                    *reinterpret_cast<uint4*>(dst + ks * ks_stride) = pt_expand16(ws[ks] & 0xFFFFu);
                    *reinterpret_cast<uint4*>(dst + ks * ks_stride + 128) = pt_expand16(ws[ks] >> 16);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the tensor core's reads
                pt_mbar_arrive(&full_bar[slot]);
            }
        }
    } else if (lane == 0) {
        // ===== MMA issuer =====
        for (int it = 0; it < n_stage; ++it) {
            const int slot = it % PT_STAGES;
            const uint32_t phase = (uint32_t)((it / PT_STAGES) & 1);
            pt_mbar_wait(&full_bar[slot], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = pt_smem_u32(stage_base + (size_t)slot * PT_STAGE_BYTES);
            const uint32_t b0 = a0 + PT_A_KSTEP * PT_KS;
#pragma unroll
            for (int ks = 0; ks < PT_KS; ++ks) {
                const uint64_t da = pt_smem_desc(a0 + ks * PT_A_KSTEP), db = pt_smem_desc(b0 + ks * PT_B_KSTEP);
                const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                    ::"r"(tmem_base), "l"(da), "l"(db), "r"(PT_IDESC), "r"(acc), "r"(0u) : "memory");
            }
            // arrives when the MMAs issued so far have read their operands: the slot is free again
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pt_smem_u32(&empty_bar[slot])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pt_smem_u32(&accum_bar)) : "memory");
    }

    if (warp < 4) {
        // ===== epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (rows of the tile), 32 columns at a time =====
        if (n_stage > 0) {
            pt_mbar_wait(&accum_bar, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t a = (int64_t)mt * PT_M + 32 * warp + lane;
#pragma unroll 1
            for (int cc = 0; cc < PT_N / 32; ++cc) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int64_t b = (int64_t)nt * PT_N + cc * 32 + j;
                    if (v[j] != 0u && a <= b && b < p) {
                        atomicAdd(pair_counts + a * p + b, (unsigned long long)v[j]);
                        if (a < b) atomicAdd(pair_counts + b * p + a, (unsigned long long)v[j]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == PT_ROWS / 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PT_N) : "memory");
}
