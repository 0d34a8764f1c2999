// probe.cu -- micro-benchmarks that put the denominators next to the numerators: bench.py runs them in
// the same process, under the same clocks, as the workload it reports ("L2 and
// SMEM peaks are not in MEASURED_PEAKS.json -- builder must measure them in the same run").
//
//   blipgpu_probe_read_bandwidth   streaming 16-byte loads over a buffer, several passes: a 48 MB
//                                  buffer stays in the 126 MB L2 (L2 -> SM bandwidth), a 4 GB one
//                                  does not (HBM read bandwidth)
//   blipgpu_probe_fp64             DFMA throughput of the fp64 pipe, and DMMA (mma.sync.m8n8k4.f64,
//                                  the instruction of the Gram build)
//   blipgpu_probe_latency          dependent-load latency of a pointer chase that hits L2
#include <algorithm>
#include "internal.cuh"

namespace {

__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ buf, size_t n_vec, int passes,
                                                   unsigned long long* __restrict__ sink)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int ps = 0; ps < passes; ++ps) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        // 8 independent 16-byte loads in flight per thread
        for (; i + 7 * stride < n_vec; i += 8 * stride) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcg(buf + i + k * stride);     // L2, not L1
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc.x ^= v[k].x; acc.y += v[k].y; acc.z ^= v[k].z; acc.w += v[k].w; }
        }
        for (; i < n_vec; i += stride) { const uint4 v = __ldcg(buf + i); acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) atomicAdd(sink, 1ull);   // keeps the loads alive
}

__global__ void __launch_bounds__(256) dfma_kernel(int iters, double seed, double* __restrict__ sink)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x + k;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) sink[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(int iters, double seed, double* __restrict__ sink)
{
    // mma.sync.aligned.m8n8k4.row.col.f64: A 8x4 (1 double / lane), B 4x8 (1 / lane), C 8x8 (2 / lane)
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = seed; c[k][1] = seed + k; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 12345.678) sink[0] = s;
}

__global__ void chase_init_kernel(unsigned* next, unsigned n, unsigned step)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) next[(size_t)i * 32] = ((i + step) % n) * 32;       // one entry per 128-byte line
}

__global__ void chase_kernel(const unsigned* __restrict__ next, int hops, unsigned* sink, long long* cycles)
{
    unsigned at = 0;
    for (int i = 0; i < 64; ++i) at = __ldcg(next + at);            // warm the path
    const long long t0 = clock64();
    for (int i = 0; i < hops; ++i) at = __ldcg(next + at);
    const long long t1 = clock64();
    *sink = at;
    *cycles = t1 - t0;
}

} // namespace

extern "C" int blipgpu_probe_read_bandwidth(blipgpu_ctx* ctx, int64_t bytes, int32_t passes, double* gb_per_s)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (bytes < 4096 || passes <= 0 || !gb_per_s) { blip_set_error("probe_read_bandwidth: bad arguments"); return BLIP_ERR_ARG; }
    DevBuf buf, sink;
    BLIP_TRY(buf.alloc((size_t)bytes));
    BLIP_TRY(sink.alloc(sizeof(unsigned long long)));
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemsetAsync(buf.ptr, 1, (size_t)bytes, st));
    CUDA_TRY(cudaMemsetAsync(sink.ptr, 0, sizeof(unsigned long long), st));
    const size_t n_vec = (size_t)bytes / 16;
    const unsigned grid = (unsigned)(ctx->sm_count * 8);
    read_kernel<<<grid, 256, 0, st>>>(buf.as<uint4>(), n_vec, 1, sink.as<unsigned long long>());   // warm-up / L2 fill
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        read_kernel<<<grid, 256, 0, st>>>(buf.as<uint4>(), n_vec, passes, sink.as<unsigned long long>());
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, (double)n_vec * 16.0 * passes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaGetLastError());
    ctx->kernel_launches += 4;
    *gb_per_s = best;
    return BLIP_OK;
}

extern "C" int blipgpu_probe_fp64(blipgpu_ctx* ctx, int32_t kind, double* tflops)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!tflops || kind < 0 || kind > 1) { blip_set_error("probe_fp64: bad arguments"); return BLIP_ERR_ARG; }
    DevBuf sink;
    BLIP_TRY(sink.alloc(sizeof(double)));
    cudaStream_t st = ctx->stream;
    const unsigned grid = (unsigned)(ctx->sm_count * 8);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                    // first repetition warms up
        CUDA_TRY(cudaEventRecord(e0, st));
        if (kind == 0) dfma_kernel<<<grid, 256, 0, st>>>(iters, 1.0, sink.as<double>());
        else dmma_kernel<<<grid, 256, 0, st>>>(iters, 1.0, sink.as<double>());
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        // DFMA: 8 fma / thread / iteration = 16 flop; DMMA: 8 mma / warp / iteration, 2*8*8*4 = 512 flop each
        const double flop = kind == 0 ? (double)grid * 256 * iters * 16.0 : (double)grid * 8 * iters * 8 * 512.0;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(cudaGetLastError());
    ctx->kernel_launches += 4;
    *tflops = best;
    return BLIP_OK;
}

extern "C" int blipgpu_probe_latency(blipgpu_ctx* ctx, int64_t bytes, double* cycles_per_load)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (bytes < (1 << 16) || !cycles_per_load) { blip_set_error("probe_latency: bad arguments"); return BLIP_ERR_ARG; }
    const unsigned n = (unsigned)std::min<int64_t>(bytes / 128, 1u << 26);
    DevBuf buf, sink, cyc;
    BLIP_TRY(buf.alloc((size_t)n * 128)); BLIP_TRY(sink.alloc(sizeof(unsigned))); BLIP_TRY(cyc.alloc(sizeof(long long)));
    cudaStream_t st = ctx->stream;
    // a stride co-prime with n visits every line once per cycle, far apart (no prefetch, no sector reuse)
    unsigned step = n / 2 + 1;
    while (std::__gcd(step, n) != 1) ++step;
    chase_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(buf.as<unsigned>(), n, step);
    const int hops = 20000;
    chase_kernel<<<1, 1, 0, st>>>(buf.as<unsigned>(), (int)std::min<unsigned>(n, hops), sink.as<unsigned>(), cyc.as<long long>());   // fills L2
    chase_kernel<<<1, 1, 0, st>>>(buf.as<unsigned>(), (int)std::min<unsigned>(n, hops), sink.as<unsigned>(), cyc.as<long long>());
    long long c = 0;
    CUDA_TRY(cudaMemcpyAsync(&c, cyc.ptr, sizeof(c), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    ctx->kernel_launches += 3;
    *cycles_per_load = (double)c / (double)std::min<unsigned>(n, hops);
    return BLIP_OK;
}
