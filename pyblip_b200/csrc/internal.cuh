// internal.cuh -- opaque handle definitions behind include/blipgpu.h
#pragma once
#include "../../include/blipgpu.h"
#include "common.cuh"
#include <cstdlib>

struct blipgpu_ctx {
    int device;
    int sm_count;
    size_t smem_optin;     // max dynamic shared memory per block (opt-in)
    size_t smem_per_sm;    // shared memory of one SM
    size_t smem_reserved;  // shared memory the system reserves per resident block
    cudaStream_t stream;   // sweep / counting kernels
    cudaStream_t stream2;  // permutation generation, copies
    double sweep_ms, other_ms;
    int64_t sweep_launches, other_launches;
    int64_t kernel_launches, h2d_bytes, d2h_bytes;
    cudaEvent_t ev[8];
    // pinned bounce buffers of blip_d2h (allocated on first use)
    void* pinned[4]; cudaEvent_t pinned_ev[4]; size_t pinned_bytes;
    // value-arena size the last sampler call of shape (chains, n_keep, p) ended up needing (gibbs.cu)
    int64_t arena_hint, arena_hint_key[3];
};

struct blipgpu_design {
    blipgpu_ctx* ctx;
    int kind;              // BLIPGPU_DESIGN_*
    int64_t n, p;
    int64_t ldx;           // row pitch of XT in doubles (multiple of 4, zero padded)
    double* XT;            // [p][ldx]  column j of X is row j here (HBM)
    double* Xl2;           // [p]       squared column norms = diag of the Gram matrix
    double* G;             // [p][ldg]  Gram matrix or NULL
    int64_t ldg;
};

struct blipgpu_samples {
    blipgpu_ctx* ctx;
    int64_t rows, rows_pad, p, n, chains, n_keep, words;
    uint32_t* bits;        // [words][rows_pad]  bit b of word w = (beta[row][32w+b] != 0)
    int64_t* row_off;      // [rows] offset of the row's values in `values`
    int32_t* row_nnz;      // [rows]
    double* values;        // arena; a row's values are stored in increasing column order
    int64_t arena_cap;
    unsigned long long* arena_used;  // device counter
    double* hyper;         // [3][rows]: p0, tau2, sigma2
    double* latents;       // [rows][n] or NULL
    int64_t n_events;
    int mode;
    double sweep_ms, total_ms;
    int64_t sweep_launches;
    int64_t n_changes, n_windows, n_sweeps_total;
};

struct blipgpu_bits {
    blipgpu_ctx* ctx;
    int64_t N, N_pad, p, words;
    uint32_t* bits;        // [words][N_pad]
    bool owns;
    // column-major copy, built on the first count_groups call that can use it:
    // colbits[c * NW + t] = bit r of column c for the rows r = 32 t .. 32 t + 31 (rows >= N are zero)
    uint32_t* colbits;     // [32 * words][NW]
    int64_t NW;            // row words per column, a multiple of 8
};

// helpers implemented in api.cu
int blip_ctx_activate(blipgpu_ctx* ctx);
// Large device -> pageable host copy: pipelined through pinned bounce buffers on ctx->stream2 and
// scattered into the destination by several host threads (first-touch page faults in parallel).
int blip_d2h(blipgpu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
// helpers implemented in gibbs.cu
int blip_fill_perms(cudaStream_t st, void* buf, int perm16, int64_t chains, int S, int64_t p,
                    int64_t sweep0, uint32_t k0, uint32_t k1, uint32_t chain_offset);
int blip_xty(cudaStream_t st, const blipgpu_design* d, const double* d_y, double* d_q0);
bool blip_host_is_pinned(const void* p);                                            // api.cu
int blip_bits_alloc(blipgpu_ctx* ctx, int64_t N, int64_t p, blipgpu_bits** out);   // counting.cu
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Device buffers of a call (sample store, sampler and counting buffers, design): stream-ordered.  A freed block of
// 1 MB or more goes to a per-stream cache of the library and is handed to the next request of (nearly) that size
// on the same stream -- stream order makes that safe -- so that a repeated job issues no driver allocation at all;
// misses and small blocks come from the device's stream-ordered pool (cudaMallocAsync, release threshold lifted).
// BLIPGPU_NO_POOL=1: plain cudaMalloc / cudaFree, no cache.  (api.cu)
static inline bool blip_use_pool() { static const bool v = getenv("BLIPGPU_NO_POOL") == nullptr; return v; }
cudaError_t blip_store_alloc(void** p, size_t bytes, cudaStream_t st);
void blip_store_free(void* p, cudaStream_t st);
void blip_store_release(cudaStream_t st);          // give a stream's idle blocks back to the driver

// device allocation freed on scope exit
// Inside a DevPoolScope the buffers come from the device's stream-ordered pool on the scope's stream (the sampler:
// ~25 buffers per call, and cudaMalloc / cudaFree were seen to take up to 100s of ms each on a busy driver).
struct DevBuf {
    void* ptr = nullptr; size_t bytes = 0; cudaStream_t pool_st = nullptr; bool pooled = false;
    static cudaStream_t& scope_stream() { static thread_local cudaStream_t st = nullptr; return st; }
    static bool& scope_on() { static thread_local bool on = false; return on; }
    int alloc(size_t b) {
        bytes = b;
        if (b == 0) { ptr = nullptr; return BLIP_OK; }
        pooled = scope_on() && blip_use_pool();
        pool_st = scope_stream();
        cudaError_t e = pooled ? blip_store_alloc(&ptr, b, pool_st) : cudaMalloc(&ptr, b);
        if (e != cudaSuccess) {
            ptr = nullptr;
            blip_set_error("device allocation of %zu bytes failed: %s", b, cudaGetErrorString(e));
            return BLIP_ERR_NOMEM;
        }
        return BLIP_OK;
    }
    ~DevBuf() {
        if (!ptr) return;
        if (pooled) blip_store_free(ptr, pool_st);
        else cudaFree(ptr);
    }
    template <class T> T* as() { return reinterpret_cast<T*>(ptr); }
};
struct DevPoolScope {
    bool prev_on; cudaStream_t prev_st;
    explicit DevPoolScope(cudaStream_t st) : prev_on(DevBuf::scope_on()), prev_st(DevBuf::scope_stream()) { DevBuf::scope_on() = true; DevBuf::scope_stream() = st; }
    ~DevPoolScope() { DevBuf::scope_on() = prev_on; DevBuf::scope_stream() = prev_st; }
};

// event-timed region on ctx->stream
struct ScopedTimer {
    blipgpu_ctx* ctx; cudaEvent_t e0, e1; bool sweep; bool ok;
    ScopedTimer(blipgpu_ctx* c, bool is_sweep) : ctx(c), sweep(is_sweep), ok(false) {
        if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
            ok = true; cudaEventRecord(e0, ctx->stream);
        }
    }
    // returns elapsed ms (synchronises on the end event)
    double stop(int64_t launches) {
        if (!ok) return 0.0;
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1); ok = false;
        if (sweep) { ctx->sweep_ms += ms; ctx->sweep_launches += launches; }
        else { ctx->other_ms += ms; ctx->other_launches += launches; }
        ctx->kernel_launches += launches;
        return ms;
    }
    ~ScopedTimer() { if (ok) { cudaEventDestroy(e0); cudaEventDestroy(e1); } }
};
