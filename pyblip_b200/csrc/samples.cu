// samples.cu -- accessors of the on-device posterior sample store.
//
// The sampler keeps (chains * n_keep) rows as bit-packed indicators plus the
// non-zero coefficients in column order.  The dense (rows, p) float64 matrix the
// reference returns (`np.concatenate([x['betas'][burn:] ...])`,
// /root/reference/pyblip/linear/linear.py:165-168) is materialised on request.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gibbs_kernels.cuh"

namespace {

// one block per row: scatter the row's values into a zero-initialised dense row
__global__ void __launch_bounds__(GB_BLOCK) dense_rows_kernel(
    const uint32_t* __restrict__ bits, const int64_t* __restrict__ row_off,
    const double* __restrict__ values, int64_t rows_pad, int words, int p, int64_t row0,
    double* __restrict__ out)
{
    __shared__ int s_scan[GB_NWARP + 2];
    const int64_t row = row0 + blockIdx.x;
    const int64_t off = row_off[row];
    double* dst = out + (int64_t)blockIdx.x * p;
    int base = 0;
    for (int w0 = 0; w0 < words; w0 += GB_BLOCK) {
        const int w = w0 + threadIdx.x;
        uint32_t m = (w < words) ? bits[(int64_t)w * rows_pad + row] : 0u;
        int total;
        int pre = block_excl_scan(__popc(m), s_scan, &total);
        const double* src = values + off + base + pre;
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            dst[32 * w + b] = *src++;
        }
        base += total;
    }
}

// CSR export: indices + values of every row, rows in order
__global__ void __launch_bounds__(GB_BLOCK) csr_rows_kernel(
    const uint32_t* __restrict__ bits, const int64_t* __restrict__ row_off,
    const double* __restrict__ values, const int64_t* __restrict__ indptr, int64_t rows_pad,
    int words, int64_t row0, int32_t* __restrict__ out_idx, double* __restrict__ out_val)
{
    __shared__ int s_scan[GB_NWARP + 2];
    const int64_t row = row0 + blockIdx.x;
    const int64_t off = row_off[row], dst0 = indptr[row];
    int base = 0;
    for (int w0 = 0; w0 < words; w0 += GB_BLOCK) {
        const int w = w0 + threadIdx.x;
        uint32_t m = (w < words) ? bits[(int64_t)w * rows_pad + row] : 0u;
        int total;
        int pre = block_excl_scan(__popc(m), s_scan, &total);
        int64_t k = base + pre;
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            out_idx[dst0 + k] = 32 * w + b;
            out_val[dst0 + k] = values[off + k];
            ++k;
        }
        base += total;
    }
}

} // namespace

extern "C" int blipgpu_samples_info_get(blipgpu_samples* s, blipgpu_samples_info* info)
{
    if (!s || !info) { blip_set_error("samples_info: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    unsigned long long used = 0;
    CUDA_TRY(cudaMemcpy(&used, s->arena_used, sizeof(used), cudaMemcpyDeviceToHost));
    info->rows = s->rows; info->p = s->p; info->n = s->n; info->chains = s->chains;
    info->n_keep = s->n_keep; info->nnz = (int64_t)used; info->n_events = s->n_events;
    info->has_latents = s->latents ? 1 : 0; info->mode = s->mode;
    info->sweep_ms = s->sweep_ms; info->sweep_launches = s->sweep_launches;
    info->total_ms = s->total_ms; info->n_changes = s->n_changes; info->n_windows = s->n_windows;
    info->n_sweeps_total = s->n_sweeps_total;
    info->bytes_device = (int64_t)(sizeof(uint32_t) * s->words * s->rows_pad + sizeof(double) * s->arena_cap +
                                   sizeof(double) * 3 * s->rows + 12 * s->rows +
                                   (s->latents ? sizeof(double) * s->rows * s->n : 0));
    return BLIP_OK;
}

extern "C" int blipgpu_samples_hyper(blipgpu_samples* s, double* p0s, double* tau2s, double* sigma2s)
{
    if (!s) { blip_set_error("samples_hyper: NULL handle"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    const size_t b = sizeof(double) * s->rows;
    s->ctx->d2h_bytes += (int64_t)b * ((p0s != nullptr) + (tau2s != nullptr) + (sigma2s != nullptr));
    if (p0s) CUDA_TRY(cudaMemcpy(p0s, s->hyper, b, cudaMemcpyDeviceToHost));
    if (tau2s) CUDA_TRY(cudaMemcpy(tau2s, s->hyper + s->rows, b, cudaMemcpyDeviceToHost));
    if (sigma2s) CUDA_TRY(cudaMemcpy(sigma2s, s->hyper + 2 * s->rows, b, cudaMemcpyDeviceToHost));
    return BLIP_OK;
}

extern "C" int blipgpu_samples_dense(blipgpu_samples* s, int64_t row0, int64_t row1, double* out)
{
    if (!s || !out || row0 < 0 || row1 > s->rows || row0 > row1) { blip_set_error("samples_dense: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    cudaStream_t st = s->ctx->stream;
    const int64_t max_rows = std::max<int64_t>(1, (int64_t)(256 << 20) / (8 * s->p));  // 256 MB staging
    const int64_t stage_rows = std::min<int64_t>(max_rows, row1 - row0);
    if (stage_rows == 0) return BLIP_OK;
    double* stage = nullptr;
    cudaError_t e = blip_store_alloc((void**)&stage, sizeof(double) * stage_rows * s->p, st);
    if (e != cudaSuccess) { blip_set_error("samples_dense: staging allocation failed: %s", cudaGetErrorString(e)); return BLIP_ERR_NOMEM; }
    int rc = BLIP_OK;
    for (int64_t r = row0; r < row1 && rc == BLIP_OK; r += stage_rows) {
        const int64_t nr = std::min<int64_t>(stage_rows, row1 - r);
        cudaMemsetAsync(stage, 0, sizeof(double) * nr * s->p, st);
        s->ctx->kernel_launches += 1; s->ctx->d2h_bytes += (int64_t)sizeof(double) * nr * s->p;
        dense_rows_kernel<<<(unsigned)nr, GB_BLOCK, 0, st>>>(s->bits, s->row_off, s->values, s->rows_pad,
                                                            (int)s->words, (int)s->p, r, stage);
        e = cudaMemcpyAsync(out + (r - row0) * s->p, stage, sizeof(double) * nr * s->p, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { blip_set_error("samples_dense: %s", cudaGetErrorString(e)); rc = BLIP_ERR_CUDA; }
    }
    blip_store_free(stage, st);
    return rc;
}

extern "C" int blipgpu_samples_csr(blipgpu_samples* s, int64_t* indptr, int32_t* indices, double* values)
{
    if (!s || !indptr) { blip_set_error("samples_csr: bad arguments"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    const bool log = getenv("BLIPGPU_LAUNCH_LOG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    std::vector<int32_t> nnz((size_t)s->rows);
    CUDA_TRY(cudaMemcpy(nnz.data(), s->row_nnz, sizeof(int32_t) * s->rows, cudaMemcpyDeviceToHost));
    indptr[0] = 0;
    for (int64_t r = 0; r < s->rows; ++r) indptr[r + 1] = indptr[r] + nnz[(size_t)r];
    const int64_t total = indptr[s->rows];
    if (!indices || !values || total == 0) return BLIP_OK;
    const double ms_indptr = ms_since(t0);
    const auto t1 = std::chrono::steady_clock::now();
    int64_t* d_indptr = nullptr; int32_t* d_idx = nullptr; double* d_val = nullptr;
    cudaStream_t st = s->ctx->stream, st2 = s->ctx->stream2;
    cudaError_t e;
    // sizes in steps of 2^24 entries: the non-zero count of a repeated job differs by a fraction of a percent from
    // call to call, and a pool that sees the same request sizes every time settles after the first call
    const int64_t cap = round_up(total, (int64_t)1 << 24);
    if ((e = blip_store_alloc((void**)&d_indptr, sizeof(int64_t) * (s->rows + 1), st)) != cudaSuccess ||
        (e = blip_store_alloc((void**)&d_idx, sizeof(int32_t) * cap, st)) != cudaSuccess ||
        (e = blip_store_alloc((void**)&d_val, sizeof(double) * cap, st)) != cudaSuccess) {
        blip_store_free(d_indptr, st); blip_store_free(d_idx, st); blip_store_free(d_val, st);
        blip_set_error("samples_csr: allocation failed: %s", cudaGetErrorString(e));
        return BLIP_ERR_NOMEM;
    }
    struct Temps { void* a; void* b; void* c; cudaStream_t st; ~Temps() { blip_store_free(a, st); blip_store_free(b, st); blip_store_free(c, st); } }
        temps{ d_indptr, d_idx, d_val, st };
    s->ctx->d2h_bytes += (int64_t)(sizeof(int32_t) + sizeof(double)) * total + (int64_t)sizeof(int32_t) * s->rows;
    s->ctx->h2d_bytes += (int64_t)sizeof(int64_t) * (s->rows + 1);
    CUDA_TRY(cudaMemcpyAsync(d_indptr, indptr, sizeof(int64_t) * (s->rows + 1), cudaMemcpyHostToDevice, st));
    const double ms_alloc = ms_since(t1);
    const auto t2 = std::chrono::steady_clock::now();
    const bool direct = blip_host_is_pinned(indices) && blip_host_is_pinned(values);
    if (!direct) {
        // pageable destination: gather everything, then the bounce-buffer copy of blip_d2h
        csr_rows_kernel<<<(unsigned)s->rows, GB_BLOCK, 0, st>>>(s->bits, s->row_off, s->values, d_indptr, s->rows_pad,
                                                               (int)s->words, 0, d_idx, d_val);
        s->ctx->kernel_launches += 1;
        CUDA_TRY(cudaGetLastError());
        BLIP_TRY(blip_d2h(s->ctx, indices, d_idx, sizeof(int32_t) * total));
        BLIP_TRY(blip_d2h(s->ctx, values, d_val, sizeof(double) * total));
        return BLIP_OK;
    }
    // page-locked destination (blipgpu_host_alloc): slabs of rows, the DMA of a slab runs under the gather of the next
    const int nslab = (int)std::min<int64_t>(8, s->rows);
    cudaEvent_t ev[8] = {};
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < nslab && err == cudaSuccess; ++k) err = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    for (int k = 0; k < nslab && err == cudaSuccess; ++k) {
        const int64_t r0 = s->rows * k / nslab, r1 = s->rows * (k + 1) / nslab;
        const int64_t a = indptr[r0], b = indptr[r1];
        csr_rows_kernel<<<(unsigned)(r1 - r0), GB_BLOCK, 0, st>>>(s->bits, s->row_off, s->values, d_indptr, s->rows_pad,
                                                                 (int)s->words, r0, d_idx, d_val);
        s->ctx->kernel_launches += 1;
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        if ((err = cudaEventRecord(ev[k], st)) != cudaSuccess) break;
        if ((err = cudaStreamWaitEvent(st2, ev[k], 0)) != cudaSuccess) break;
        if (b > a) {
            if ((err = cudaMemcpyAsync(indices + a, d_idx + a, sizeof(int32_t) * (b - a), cudaMemcpyDeviceToHost, st2)) != cudaSuccess) break;
            if ((err = cudaMemcpyAsync(values + a, d_val + a, sizeof(double) * (b - a), cudaMemcpyDeviceToHost, st2)) != cudaSuccess) break;
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(st2);       // the temporaries are freed on `st` after this
    if (err == cudaSuccess) err = e2;
    e2 = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = e2;
    for (int k = 0; k < nslab; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
    if (err != cudaSuccess) { blip_set_error("samples_csr: %s", cudaGetErrorString(err)); return BLIP_ERR_CUDA; }
    if (log) fprintf(stderr, "[blipgpu] samples_csr: row counts + indptr %.1f ms, temporaries %.1f ms, %d slabs of gather + DMA %.1f ms (host clock)\n",
                     ms_indptr, ms_alloc, nslab, ms_since(t2));
    return BLIP_OK;
}

extern "C" int blipgpu_samples_latents(blipgpu_samples* s, int64_t row0, int64_t row1, double* out)
{
    if (!s || !out || row0 < 0 || row1 > s->rows || row0 > row1) { blip_set_error("samples_latents: bad arguments"); return BLIP_ERR_ARG; }
    if (!s->latents) { blip_set_error("samples_latents: no latents were kept (not a probit run?)"); return BLIP_ERR_ARG; }
    BLIP_TRY(blip_ctx_activate(s->ctx));
    CUDA_TRY(cudaMemcpy(out, s->latents + row0 * s->n, sizeof(double) * (row1 - row0) * s->n, cudaMemcpyDeviceToHost));
    return BLIP_OK;
}

extern "C" int blipgpu_samples_free(blipgpu_samples* s)
{
    if (!s) return BLIP_OK;
    cudaSetDevice(s->ctx->device);
    // stream-ordered frees: the memory returns to the device pool (kept there, blipgpu_ctx_create) once everything
    // queued on the library's streams has used it
    cudaStreamSynchronize(s->ctx->stream2);
    cudaStream_t st = s->ctx->stream;
    void* ptrs[] = { s->bits, s->row_off, s->row_nnz, s->values, s->arena_used, s->hyper, s->latents };
    for (void* ptr : ptrs) blip_store_free(ptr, st);
    delete s;
    return BLIP_OK;
}
