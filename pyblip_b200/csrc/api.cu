// api.cu -- context, error reporting and design-matrix handling of libblipgpu.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <algorithm>
#include <cstring>
#include <sched.h>
#include <cstdlib>
#include <sys/mman.h>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
#include "internal.cuh"

static thread_local char g_err[1024] = "";

void blip_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int blipgpu_abi_version(void) { return BLIPGPU_ABI_VERSION; }
extern "C" const char* blipgpu_last_error(void) { return g_err; }
extern "C" int blipgpu_sizeof(int which)
{
    switch (which) {
        case 0: return (int)sizeof(blipgpu_spikeslab_config);
        case 1: return (int)sizeof(blipgpu_streams);
        case 2: return (int)sizeof(blipgpu_sample_options);
        case 3: return (int)sizeof(blipgpu_samples_info);
        case 4: return (int)sizeof(blipgpu_multi_streams);
        case 5: return (int)sizeof(blipgpu_nprior_config);
        case 6: return (int)sizeof(blipgpu_nprior_streams);
        default: return -1;
    }
}

extern "C" int blipgpu_device_count(int* count)
{
    if (!count) { blip_set_error("count is NULL"); return BLIP_ERR_ARG; }
    CUDA_TRY(cudaGetDeviceCount(count));
    return BLIP_OK;
}

int blip_ctx_activate(blipgpu_ctx* ctx)
{
    if (!ctx) { blip_set_error("NULL context"); return BLIP_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    return BLIP_OK;
}

// ---- device blocks, cached per stream (internal.cuh) ------------------------------------------------------
// The driver's stream-ordered pool alone was not enough: with identical request sequences it still went to the
// driver for fresh memory now and then (0.08 - 1.1 s inside the first timed sampler call of a bench run, 0.5 s in a
// samples_csr()), which a size-keyed free list in front of it does not.
namespace {
struct DevBlock { void* ptr; size_t bytes; };
struct DevCache { std::vector<DevBlock> idle; size_t idle_bytes = 0; };        // oldest first
struct DevCaches {
    std::mutex mu;
    std::unordered_map<cudaStream_t, DevCache> by_stream;
    std::unordered_map<void*, size_t> live;
};
DevCaches& dev_caches() { static DevCaches* c = new DevCaches(); return *c; }
constexpr size_t kDevCacheMin = (size_t)1 << 20;
size_t dev_cache_limit()
{
    static const size_t v = [] {
        const char* e = getenv("BLIPGPU_DEVICE_CACHE_MB");
        return (size_t)(e ? std::max(0, atoi(e)) : 65536) << 20;
    }();
    return v;
}
cudaError_t raw_dev_alloc(void** p, size_t bytes, cudaStream_t st)
{
    return blip_use_pool() ? cudaMallocAsync(p, bytes, st) : cudaMalloc(p, bytes);
}
void raw_dev_free(void* p, cudaStream_t st)
{
    if (blip_use_pool() && cudaFreeAsync(p, st) == cudaSuccess) return;
    cudaGetLastError();
    cudaStreamSynchronize(st);
    cudaFree(p);
}
} // namespace

cudaError_t blip_store_alloc(void** p, size_t bytes, cudaStream_t st)
{
    if (bytes < kDevCacheMin || !blip_use_pool()) return raw_dev_alloc(p, bytes, st);
    const size_t want = (size_t)round_up((int64_t)bytes, (int64_t)2 << 20);
    DevCaches& dc = dev_caches();
    {
        std::lock_guard<std::mutex> lock(dc.mu);
        DevCache& c = dc.by_stream[st];
        int best = -1;
        for (int i = 0; i < (int)c.idle.size(); ++i)      // best fit, at most a quarter (+ 8 MB) larger than asked
            if (c.idle[i].bytes >= want && c.idle[i].bytes <= want + want / 4 + ((size_t)8 << 20) &&
                (best < 0 || c.idle[i].bytes < c.idle[best].bytes)) best = i;
        if (best >= 0) {
            const DevBlock b = c.idle[best];
            c.idle.erase(c.idle.begin() + best);
            c.idle_bytes -= b.bytes;
            dc.live[b.ptr] = b.bytes;
            *p = b.ptr;
            return cudaSuccess;
        }
    }
    cudaError_t e = raw_dev_alloc(p, want, st);
    if (e != cudaSuccess) {                                // out of memory: give every idle block back and try again
        cudaGetLastError();
        std::vector<std::pair<cudaStream_t, DevBlock>> drop;
        {
            std::lock_guard<std::mutex> lock(dc.mu);
            for (auto& kv : dc.by_stream) {
                for (const DevBlock& b : kv.second.idle) drop.push_back({ kv.first, b });
                kv.second.idle.clear(); kv.second.idle_bytes = 0;
            }
        }
        for (auto& d : drop) raw_dev_free(d.second.ptr, d.first);
        cudaStreamSynchronize(st);
        e = raw_dev_alloc(p, want, st);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(dc.mu);
        dc.live[*p] = want;
    }
    return e;
}

void blip_store_free(void* p, cudaStream_t st)
{
    if (!p) return;
    DevCaches& dc = dev_caches();
    std::vector<DevBlock> drop;
    {
        std::lock_guard<std::mutex> lock(dc.mu);
        auto it = dc.live.find(p);
        if (it == dc.live.end()) drop.push_back({ p, 0 });                      // a small block, or from before the cache
        else {
            DevCache& c = dc.by_stream[st];
            c.idle.push_back({ p, it->second });
            c.idle_bytes += it->second;
            dc.live.erase(it);
            while (c.idle_bytes > dev_cache_limit() && !c.idle.empty()) {
                drop.push_back(c.idle.front());
                c.idle_bytes -= c.idle.front().bytes;
                c.idle.erase(c.idle.begin());
            }
        }
    }
    for (const DevBlock& b : drop) raw_dev_free(b.ptr, st);
}

void blip_store_release(cudaStream_t st)
{
    DevCaches& dc = dev_caches();
    std::vector<DevBlock> drop;
    {
        std::lock_guard<std::mutex> lock(dc.mu);
        auto it = dc.by_stream.find(st);
        if (it == dc.by_stream.end()) return;
        drop.swap(it->second.idle);
        dc.by_stream.erase(it);
    }
    for (const DevBlock& b : drop) raw_dev_free(b.ptr, st);
}

// ---- page-locked host blocks for results, kept across calls --------------------------------------------
// A result that goes into page-locked memory is one DMA; into a fresh pageable array it is a bounce through
// a pinned buffer plus a host copy that takes a page fault per page (0.12 - 0.3 s for the 1.6 GB CSR of C2).
// Pinning itself is slow (~0.25 ms per MB), so blocks are recycled: the host side wraps a block in a NumPy
// array whose finaliser hands it back (pyblip_b200/_lib.py: host_empty).  Process-wide, not per context: a
// finaliser may run after the context is gone.
namespace {
struct HostBlock { void* ptr; size_t bytes; int device; };
struct HostCache {
    std::mutex mu;
    std::unordered_map<void*, HostBlock> live;
    std::vector<HostBlock> idle;                    // oldest first
    size_t idle_bytes = 0;
};
HostCache& host_cache() { static HostCache* c = new HostCache(); return *c; }   // never destroyed (finalisers at exit)
size_t host_cache_limit()
{
    static const size_t v = [] {
        const char* e = getenv("BLIPGPU_HOST_CACHE_MB");
        return (size_t)(e ? std::max(0, atoi(e)) : 8192) << 20;
    }();
    return v;
}
void host_block_release(const HostBlock& b)
{
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != b.device) cudaSetDevice(b.device);
    cudaFreeHost(b.ptr);
    if (cur != b.device) cudaSetDevice(cur);
    cudaGetLastError();
}
} // namespace

extern "C" int blipgpu_host_alloc(size_t bytes, void** out)
{
    if (!out) { blip_set_error("host_alloc: out is NULL"); return BLIP_ERR_ARG; }
    const size_t want = (size_t)round_up((int64_t)std::max<size_t>(bytes, 1), (int64_t)2 << 20);
    HostCache& hc = host_cache();
    std::lock_guard<std::mutex> lock(hc.mu);
    int best = -1;
    for (int i = 0; i < (int)hc.idle.size(); ++i)   // best fit, and no block more than twice the request
        if (hc.idle[i].bytes >= want && hc.idle[i].bytes <= 2 * want + ((size_t)8 << 20) &&
            (best < 0 || hc.idle[i].bytes < hc.idle[best].bytes)) best = i;
    if (best >= 0) {
        HostBlock b = hc.idle[best];
        hc.idle.erase(hc.idle.begin() + best);
        hc.idle_bytes -= b.bytes;
        hc.live[b.ptr] = b;
        *out = b.ptr;
        return BLIP_OK;
    }
    HostBlock b{ nullptr, want + want / 16, 0 };    // head-room: the next job's result is rarely the same size to the byte
    b.bytes = (size_t)round_up((int64_t)b.bytes, (int64_t)2 << 20);
    cudaGetDevice(&b.device);
    cudaError_t e = cudaHostAlloc(&b.ptr, b.bytes, cudaHostAllocPortable);
    if (e != cudaSuccess && !hc.idle.empty()) {     // give the idle blocks back and try once more
        cudaGetLastError();
        for (const HostBlock& i : hc.idle) host_block_release(i);
        hc.idle.clear(); hc.idle_bytes = 0;
        e = cudaHostAlloc(&b.ptr, b.bytes, cudaHostAllocPortable);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        blip_set_error("host_alloc: cannot page-lock %zu bytes: %s", b.bytes, cudaGetErrorString(e));
        return BLIP_ERR_NOMEM;
    }
    hc.live[b.ptr] = b;
    *out = b.ptr;
    return BLIP_OK;
}

extern "C" int blipgpu_host_free(void* p)
{
    if (!p) return BLIP_OK;
    HostCache& hc = host_cache();
    std::lock_guard<std::mutex> lock(hc.mu);
    auto it = hc.live.find(p);
    if (it == hc.live.end()) { blip_set_error("host_free: not a block of blipgpu_host_alloc"); return BLIP_ERR_ARG; }
    const HostBlock b = it->second;
    hc.live.erase(it);
    hc.idle.push_back(b);
    hc.idle_bytes += b.bytes;
    while (hc.idle_bytes > host_cache_limit() && !hc.idle.empty()) {
        const HostBlock old = hc.idle.front();
        hc.idle.erase(hc.idle.begin());
        hc.idle_bytes -= old.bytes;
        host_block_release(old);
    }
    return BLIP_OK;
}

bool blip_host_is_pinned(const void* p)
{
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

int blip_d2h(blipgpu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes)
{
    const size_t kChunk = (size_t)32 << 20;
    if (bytes < 2 * kChunk || blip_host_is_pinned(dst_host)) {   // small (the driver's own staging is fine) or one DMA
        CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return BLIP_OK;
    }
    {
        // the destination is usually a fresh anonymous mapping (np.empty): ask for huge pages, so that
        // the scatter below takes one page fault per 2 MB instead of per 4 KB (no-op where THP is off)
        const uintptr_t huge = (uintptr_t)2 << 20;
        const uintptr_t a = ((uintptr_t)dst_host + huge - 1) & ~(huge - 1), e = ((uintptr_t)dst_host + bytes) & ~(huge - 1);
        if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
    }
    if (!ctx->pinned[0]) {
        for (int i = 0; i < 4; ++i) {
            CUDA_TRY(cudaHostAlloc(&ctx->pinned[i], kChunk, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->pinned_ev[i], cudaEventDisableTiming));
        }
        ctx->pinned_bytes = kChunk;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the source is produced on the main stream
    const size_t nchunk = (bytes + kChunk - 1) / kChunk;
    // host threads of the scatter: the cores this process may use (sched affinity), shared with the other
    // ranks of the box when torchrun started several (LOCAL_WORLD_SIZE): eight ranks with eight threads
    // each oversubscribed a 32-core box and cost 14 % of the end-to-end figure at N = 8
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    {
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = std::max(1, CPU_COUNT(&set));
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? std::max(1, atoi(lws)) : 1;
        hw = std::max(1u, hw / (unsigned)ranks);
    }
    const unsigned nthread = std::min(8u, hw);
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
        cudaError_t e = cudaMemcpyAsync(ctx->pinned[c & 3], (const char*)src_dev + off, len, cudaMemcpyDeviceToHost, ctx->stream2);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->pinned_ev[c & 3], ctx->stream2);
        return e;
    };
    for (size_t c = 0; c < std::min<size_t>(3, nchunk); ++c) CUDA_TRY(issue(c));
    for (size_t c = 0; c < nchunk; ++c) {
        if (c + 3 < nchunk) CUDA_TRY(issue(c + 3)); // buffer (c+3)&3 was drained in the previous iteration
        CUDA_TRY(cudaEventSynchronize(ctx->pinned_ev[c & 3]));
        const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
        const char* src = (const char*)ctx->pinned[c & 3];
        char* dst = (char*)dst_host + off;
        const size_t part = ((len + nthread - 1) / nthread + 4095) & ~(size_t)4095;
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthread; ++t) {
            const size_t a = (size_t)t * part;
            if (a >= len) break;
            th.emplace_back([=] { memcpy(dst + a, src + a, std::min(part, len - a)); });
        }
        memcpy(dst, src, std::min(part, len));
        for (auto& t : th) t.join();
    }
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_create(int device, blipgpu_ctx** out)
{
    if (!out) { blip_set_error("out is NULL"); return BLIP_ERR_ARG; }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        blip_set_error("device %d out of range (%d CUDA devices visible)", device, ndev);
        return BLIP_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        blip_set_error("libblipgpu is built for sm_100a (B200); device %d is sm_%d%d", device,
                       prop.major, prop.minor);
        return BLIP_ERR_UNSUPPORTED;
    }
    blipgpu_ctx* ctx = new (std::nothrow) blipgpu_ctx();
    if (!ctx) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    ctx->smem_reserved = prop.reservedSharedMemPerBlock;
    ctx->sweep_ms = ctx->other_ms = 0.0;
    ctx->sweep_launches = ctx->other_launches = 0;
    ctx->kernel_launches = ctx->h2d_bytes = ctx->d2h_bytes = 0;
    for (int i = 0; i < 8; ++i) CUDA_TRY(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i < 4; ++i) { ctx->pinned[i] = nullptr; ctx->pinned_ev[i] = nullptr; }
    ctx->pinned_bytes = 0;
    ctx->arena_hint = 0; ctx->arena_hint_key[0] = ctx->arena_hint_key[1] = ctx->arena_hint_key[2] = 0;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    {
        // The sample store (GBs per sampler call) comes from the device's stream-ordered pool and goes back to it:
        // keep freed memory in the pool, so that a process that samples repeatedly pays the driver's
        // physical allocation once (cudaMalloc / cudaFree of GB-sized buffers were seen to take 100s of ms)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_destroy(blipgpu_ctx* ctx)
{
    if (!ctx) return BLIP_OK;
    cudaSetDevice(ctx->device);
    blip_store_release(ctx->stream);
    blip_store_release(ctx->stream2);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 4; ++i) { if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]); if (ctx->pinned_ev[i]) cudaEventDestroy(ctx->pinned_ev[i]); }
    delete ctx;
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_synchronize(blipgpu_ctx* ctx)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream2));
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_timing(blipgpu_ctx* ctx, double* sweep_ms, int64_t* sweep_launches,
                                  double* other_ms, int64_t* other_launches, int reset)
{
    if (!ctx) { blip_set_error("NULL context"); return BLIP_ERR_ARG; }
    if (sweep_ms) *sweep_ms = ctx->sweep_ms;
    if (sweep_launches) *sweep_launches = ctx->sweep_launches;
    if (other_ms) *other_ms = ctx->other_ms;
    if (other_launches) *other_launches = ctx->other_launches;
    if (reset) { ctx->sweep_ms = ctx->other_ms = 0.0; ctx->sweep_launches = ctx->other_launches = 0; }
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_counters(blipgpu_ctx* ctx, int64_t* kernel_launches, int64_t* h2d_bytes,
                                    int64_t* d2h_bytes, int reset)
{
    if (!ctx) { blip_set_error("NULL context"); return BLIP_ERR_ARG; }
    if (kernel_launches) *kernel_launches = ctx->kernel_launches;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    if (reset) ctx->kernel_launches = ctx->h2d_bytes = ctx->d2h_bytes = 0;
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_event_record(blipgpu_ctx* ctx, int slot)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (slot < 0 || slot >= 8) { blip_set_error("event slot out of range"); return BLIP_ERR_ARG; }
    CUDA_TRY(cudaEventRecord(ctx->ev[slot], ctx->stream));
    return BLIP_OK;
}

extern "C" int blipgpu_ctx_event_elapsed_ms(blipgpu_ctx* ctx, int slot_begin, int slot_end, double* ms)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (slot_begin < 0 || slot_begin >= 8 || slot_end < 0 || slot_end >= 8 || !ms) {
        blip_set_error("event_elapsed: bad arguments"); return BLIP_ERR_ARG;
    }
    CUDA_TRY(cudaEventSynchronize(ctx->ev[slot_end]));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev[slot_begin], ctx->ev[slot_end]));
    *ms = f;
    return BLIP_OK;
}

// ---------------------------------------------------------------------------
// design upload: X (n x p row-major) -> XT (p x ldx), Xl2
// ---------------------------------------------------------------------------
// 32x32 tile transpose through shared memory; coalesced on both sides.
__global__ void transpose_kernel(const double* __restrict__ X, double* __restrict__ XT, int64_t n,
                                 int64_t p, int64_t ldx)
{
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < p) ? X[i * p + j] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t j = j0 + r, i = i0 + threadIdx.x;
        if (j < p && i < ldx) XT[j * ldx + i] = tile[threadIdx.x][r];
    }
}

// one warp per column: Xl2[j] = sum_i XT[j][i]^2
__global__ void xl2_kernel(const double* __restrict__ XT, double* __restrict__ Xl2, int64_t n,
                           int64_t p, int64_t ldx)
{
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p) return;
    const double* col = XT + j * ldx;
    double s = 0.0;
    for (int64_t i = threadIdx.x & 31; i < n; i += 32) s = fma(col[i], col[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) Xl2[j] = s;
}

extern "C" int blipgpu_design_upload(blipgpu_ctx* ctx, const double* X, int64_t n, int64_t p,
                                     blipgpu_design** out)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!X || !out || n <= 0 || p <= 0) { blip_set_error("design_upload: bad arguments"); return BLIP_ERR_ARG; }
    if (p >= (int64_t)1 << 31 || n >= (int64_t)1 << 31) {
        blip_set_error("design_upload: n and p must be < 2^31"); return BLIP_ERR_ARG;
    }
    blipgpu_design* d = new (std::nothrow) blipgpu_design();
    if (!d) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    d->ctx = ctx; d->kind = BLIPGPU_DESIGN_DENSE; d->n = n; d->p = p;
    d->ldx = round_up(n, 4); d->G = nullptr; d->ldg = 0; d->XT = nullptr; d->Xl2 = nullptr;
    struct DesignGuard { blipgpu_design* d; ~DesignGuard() { if (d) blipgpu_design_free(d); } } guard{ d };
    DevPoolScope pool_scope(ctx->stream);          // design buffers: stream-ordered pool, like the sampler's (internal.cuh)
    DevBuf Xrow;                                   // row-major staging copy, freed on every exit
    BLIP_TRY(Xrow.alloc(sizeof(double) * n * p));
    cudaError_t e;
    if ((e = blip_store_alloc((void**)&d->XT, sizeof(double) * p * d->ldx, ctx->stream)) != cudaSuccess ||
        (e = blip_store_alloc((void**)&d->Xl2, sizeof(double) * p, ctx->stream)) != cudaSuccess) {
        blip_set_error("design_upload: device allocation failed: %s", cudaGetErrorString(e));
        return BLIP_ERR_NOMEM;
    }
    CUDA_TRY(cudaMemcpyAsync(Xrow.ptr, X, sizeof(double) * n * p, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((p + 31) / 32), (unsigned)((d->ldx + 31) / 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(Xrow.as<double>(), d->XT, n, p, d->ldx);
    xl2_kernel<<<(unsigned)((p + 7) / 8), 256, 0, ctx->stream>>>(d->XT, d->Xl2, n, p, d->ldx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->kernel_launches += 2; ctx->h2d_bytes += (int64_t)sizeof(double) * n * p;
    guard.d = nullptr;
    *out = d;
    return BLIP_OK;
}

__global__ void changepoint_xl2_kernel(double* Xl2, int64_t T)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < T) Xl2[j] = (double)(T - j);   // column j has T-j ones
}

extern "C" int blipgpu_design_changepoint(blipgpu_ctx* ctx, int64_t T, blipgpu_design** out)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (!out || T <= 0 || T >= (int64_t)1 << 31) { blip_set_error("design_changepoint: bad T"); return BLIP_ERR_ARG; }
    blipgpu_design* d = new (std::nothrow) blipgpu_design();
    if (!d) { blip_set_error("host allocation failed"); return BLIP_ERR_NOMEM; }
    d->ctx = ctx; d->kind = BLIPGPU_DESIGN_CHANGEPOINT; d->n = T; d->p = T; d->ldx = 0;
    d->XT = nullptr; d->G = nullptr; d->ldg = 0; d->Xl2 = nullptr;
    struct DesignGuard { blipgpu_design* d; ~DesignGuard() { if (d) blipgpu_design_free(d); } } guard{ d };
    CUDA_TRY(blip_store_alloc((void**)&d->Xl2, sizeof(double) * T, ctx->stream));
    changepoint_xl2_kernel<<<(unsigned)((T + 255) / 256), 256, 0, ctx->stream>>>(d->Xl2, T);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    guard.d = nullptr;
    *out = d;
    return BLIP_OK;
}

int blip_gram_build(blipgpu_design* d);   // gram.cu

extern "C" int blipgpu_design_build_gram(blipgpu_design* d)
{
    if (!d) { blip_set_error("NULL design"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(d->ctx));
    if (d->kind == BLIPGPU_DESIGN_CHANGEPOINT) return BLIP_OK;  // closed form, nothing stored
    if (d->G) return BLIP_OK;
    return blip_gram_build(d);
}

extern "C" int blipgpu_design_get_gram(blipgpu_design* d, int64_t r0, int64_t r1, double* out)
{
    if (!d || !out || r0 < 0 || r1 > d->p || r0 > r1) { blip_set_error("get_gram: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(d->ctx));
    if (!d->G) { blip_set_error("get_gram: Gram matrix has not been built"); return BLIP_ERR_ARG; }
    CUDA_TRY(cudaMemcpy2D(out, sizeof(double) * d->p, d->G + r0 * d->ldg, sizeof(double) * d->ldg,
                          sizeof(double) * d->p, (size_t)(r1 - r0), cudaMemcpyDeviceToHost));
    return BLIP_OK;
}

extern "C" int blipgpu_design_get_xl2(blipgpu_design* d, double* out)
{
    if (!d || !out) { blip_set_error("get_xl2: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(d->ctx));
    CUDA_TRY(cudaMemcpy(out, d->Xl2, sizeof(double) * d->p, cudaMemcpyDeviceToHost));
    return BLIP_OK;
}

extern "C" int blipgpu_design_free(blipgpu_design* d)
{
    if (!d) return BLIP_OK;
    cudaSetDevice(d->ctx->device);
    // stream-ordered frees on the stream every user of these buffers ran on (the sampler joins its second
    // stream before it returns); a cudaFree of the 0.8 GB Gram matrix synchronises the device and was seen to
    // take > 100 ms
    cudaStream_t st = d->ctx->stream;
    blip_store_free(d->XT, st); blip_store_free(d->Xl2, st); blip_store_free(d->G, st);
    delete d;
    return BLIP_OK;
}

// ---------------------------------------------------------------------------
// value-level pin of the truncated normal: the device arithmetic of common.cuh on recorded
// variate sequences (see include/blipgpu.h)
// ---------------------------------------------------------------------------
__global__ void debug_truncnorm_kernel(int64_t count, const double* mean, const double* var, const double* b,
                                       const int32_t* lower, const double* u, int64_t nu, const int64_t* u_at,
                                       const double* z, int64_t nz, const int64_t* z_at, double* out,
                                       int64_t* u_used, int64_t* z_used)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    TnInjected src{ u, z, u_at[i], z_at[i], nu, nz };
    out[i] = truncnorm_from(src, mean[i], var[i], b[i], lower[i]);
    u_used[i] = src.iu - u_at[i];
    z_used[i] = src.iz - z_at[i];
}

extern "C" int blipgpu_debug_truncnorm(blipgpu_ctx* ctx, int64_t count, const double* mean, const double* var,
                                       const double* b, const int32_t* lower_interval, const double* u,
                                       int64_t nu, const int64_t* u_at, const double* z, int64_t nz,
                                       const int64_t* z_at, double* out, int64_t* u_used, int64_t* z_used)
{
    BLIP_TRY(blip_ctx_activate(ctx));
    if (count <= 0 || !mean || !var || !b || !lower_interval || !u_at || !z_at || !out || !u_used || !z_used) {
        blip_set_error("debug_truncnorm: bad arguments"); return BLIP_ERR_ARG;
    }
    DevBuf d_mean, d_var, d_b, d_lo, d_u, d_uat, d_z, d_zat, d_out, d_uu, d_zu;
    auto up = [&](DevBuf& buf, const void* src, size_t bytes) -> int {
        BLIP_TRY(buf.alloc(std::max<size_t>(bytes, 8)));
        if (bytes) CUDA_TRY(cudaMemcpyAsync(buf.ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return BLIP_OK;
    };
    BLIP_TRY(up(d_mean, mean, sizeof(double) * count)); BLIP_TRY(up(d_var, var, sizeof(double) * count));
    BLIP_TRY(up(d_b, b, sizeof(double) * count)); BLIP_TRY(up(d_lo, lower_interval, sizeof(int32_t) * count));
    BLIP_TRY(up(d_u, u, sizeof(double) * nu)); BLIP_TRY(up(d_uat, u_at, sizeof(int64_t) * count));
    BLIP_TRY(up(d_z, z, sizeof(double) * nz)); BLIP_TRY(up(d_zat, z_at, sizeof(int64_t) * count));
    BLIP_TRY(d_out.alloc(sizeof(double) * count)); BLIP_TRY(d_uu.alloc(sizeof(int64_t) * count));
    BLIP_TRY(d_zu.alloc(sizeof(int64_t) * count));
    debug_truncnorm_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
        count, d_mean.as<double>(), d_var.as<double>(), d_b.as<double>(), d_lo.as<int32_t>(), d_u.as<double>(), nu,
        d_uat.as<int64_t>(), d_z.as<double>(), nz, d_zat.as<int64_t>(), d_out.as<double>(), d_uu.as<int64_t>(),
        d_zu.as<int64_t>());
    CUDA_TRY(cudaGetLastError());
    ctx->kernel_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(out, d_out.ptr, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(u_used, d_uu.ptr, sizeof(int64_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(z_used, d_zu.ptr, sizeof(int64_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BLIP_OK;
}
