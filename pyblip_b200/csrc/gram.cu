// gram.cu -- one-off Gram build G = X^T X in fp64 on the tensor cores.
//
// Replaces `XTX = np.dot(X.T, X)` (/root/reference/pyblip/linear/_linear_multi.pyx:379)
// and feeds the Gram-mode sweep (state q = X^T r).  fp64 has no tcgen05 kind, so
// the tensor path for doubles on sm_100a is DMMA (mma.sync.m8n8k4.f64); operands
// are staged with cp.async (LDGSTS) into a double-buffered, bank-conflict-free
// shared-memory tile.  Only tiles on or above the diagonal are computed; the
// mirror image is written from the same accumulators so G is exactly symmetric.
#include "internal.cuh"

namespace {

constexpr int TILE = 128;          // CTA computes a TILE x TILE block of G
constexpr int KT = 16;             // K (= observations) per pipeline stage
constexpr int LDS = KT + 4;        // smem row pitch in doubles: (4*row + k) mod 16 distinct
constexpr int THREADS = 512;       // 16 warps, 4x4, warp tile 32x32

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;       // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// XT: [p][ldx] (row j = column j of X, zero padded to ldx).  G: [p][ldg].
__global__ void __launch_bounds__(THREADS, 1)
gram_dmma_kernel(const double* __restrict__ XT, double* __restrict__ G, int p, int n, int64_t ldx,
                 int64_t ldg, int ntile)
{
    extern __shared__ __align__(16) double smem[];
    // upper-triangular tile index -> (bi, bj), bi <= bj
    int t = blockIdx.x, bi = 0;
    while (t >= ntile - bi) { t -= ntile - bi; ++bi; }
    const int bj = bi + t;
    const int row0 = bi * TILE, col0 = bj * TILE;

    double* sA[2] = { smem, smem + 2 * TILE * LDS };
    double* sB[2] = { smem + TILE * LDS, smem + 3 * TILE * LDS };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;   // warp tile origin inside the CTA tile

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = (n + KT - 1) / KT;
    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * KT;
        // 2 * TILE rows x (KT/2) 16-byte chunks
        for (int c = tid; c < 2 * TILE * (KT / 2); c += THREADS) {
            int r = c / (KT / 2), kc = (c % (KT / 2)) * 2;
            bool isB = r >= TILE;
            int rr = isB ? r - TILE : r;
            int grow = (isB ? col0 : row0) + rr;
            bool valid = grow < p && (k0 + kc) < ldx;
            const double* src = XT + (int64_t)(valid ? grow : 0) * ldx + (valid ? k0 + kc : 0);
            double* dst = (isB ? sB[stage] : sA[stage]) + rr * LDS + kc;
            cp_async16(dst, src, valid);
        }
        cp_async_commit();
    };

    load_stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) { load_stage(cur ^ 1, kt + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* A = sA[cur] + (wm + (lane >> 2)) * LDS + (lane & 3);
        const double* B = sB[cur] + (wn + (lane >> 2)) * LDS + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = A[i * 8 * LDS + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = B[j * 8 * LDS + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    // epilogue: C fragment lane l holds (row l/4, cols 2*(l%4), 2*(l%4)+1)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = row0 + wm + i * 8 + (lane >> 2);
            int c = col0 + wn + j * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int cc = c + e;
                if (r < p && cc < p) {
                    double v = acc[i][j][e];
                    if (bi == bj) {
                        // diagonal tile: keep exact symmetry by writing only r <= cc and its mirror
                        if (r <= cc) { G[(int64_t)r * ldg + cc] = v; G[(int64_t)cc * ldg + r] = v; }
                    } else {
                        G[(int64_t)r * ldg + cc] = v;
                        G[(int64_t)cc * ldg + r] = v;
                    }
                }
            }
        }
}

} // namespace

int blip_gram_build(blipgpu_design* d)
{
    blipgpu_ctx* ctx = d->ctx;
    const int64_t p = d->p;
    d->ldg = round_up(p, 4);
    cudaError_t e = blip_store_alloc((void**)&d->G, sizeof(double) * (size_t)p * d->ldg, ctx->stream);
    if (e != cudaSuccess) {
        d->G = nullptr;
        blip_set_error("gram build: cannot allocate %.2f GB for the %lld x %lld Gram matrix: %s",
                       8.0 * p * d->ldg / 1e9, (long long)p, (long long)p, cudaGetErrorString(e));
        return BLIP_ERR_NOMEM;
    }
    // anything that fails below must not leave a half-built matrix behind: a later
    // blipgpu_design_build_gram would take a non-NULL G for a finished one
    struct GramGuard { blipgpu_design* d; bool ok; ~GramGuard() { if (!ok) { blip_store_free(d->G, d->ctx->stream); d->G = nullptr; } } } guard{ d, false };
    CUDA_TRY(cudaMemsetAsync(d->G, 0, sizeof(double) * (size_t)p * d->ldg, ctx->stream));
    const int ntile = (int)((p + TILE - 1) / TILE);
    const int64_t nblocks = (int64_t)ntile * (ntile + 1) / 2;
    const size_t smem = sizeof(double) * 4 * TILE * LDS;
    CUDA_TRY(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScopedTimer timer(ctx, false);
    gram_dmma_kernel<<<(unsigned)nblocks, THREADS, smem, ctx->stream>>>(d->XT, d->G, (int)p, (int)d->n,
                                                                       d->ldx, d->ldg, ntile);
    CUDA_TRY(cudaGetLastError());
    timer.stop(1);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    guard.ok = true;
    return BLIP_OK;
}
