// FP64 pipe probe for sm_100a: dependent-chain latency and per-SM throughput of DFMA, DMUL,
// 64-bit shuffles and double division.   nvcc -arch=sm_100a -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double* out, long long* cyc, int iters)
{
    double a = out[0], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_ffma(float* out, long long* cyc, int iters)
{
    float a = out[0], b = 1.0000001f, c = 1e-9f;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fmaf(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_shfl(double* out, long long* cyc, int iters)
{
    double a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a = __shfl_sync(0xffffffffu, a, (j + i) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_div(double* out, long long* cyc, int iters)
{
    double a = out[0] + 3.0, b = 1.0000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a = b / a + 1.5;
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lat_log(double* out, long long* cyc, int iters)
{
    double a = out[0] + 3.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a = log(a + 2.5) + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// throughput: 8 independent chains per thread
__global__ void thr_dfma(double* out, long long* cyc, int iters)
{
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = out[k] + k;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double* d; long long* c; float* f;
    cudaMalloc(&d, 8 << 20); cudaMalloc(&f, 4 << 20); cudaMalloc(&c, 8);
    cudaMemset(d, 0, 8 << 20); cudaMemset(f, 0, 4 << 20);
    long long h; const int iters = 2000;
    auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %8.1f cycles per op\n", name, (double)h / per); };
    lat_dfma<<<1, 32>>>(d, c, iters); rep("DFMA dependent chain, 1 warp", iters * 16.0);
    lat_ffma<<<1, 32>>>(f, c, iters); rep("FFMA dependent chain, 1 warp", iters * 16.0);
    lat_shfl<<<1, 32>>>(d, c, iters); rep("64-bit shuffle dependent chain, 1 warp", iters * 16.0);
    lat_div<<<1, 32>>>(d, c, iters); rep("double division (+add) dependent, 1 warp", iters * 16.0);
    lat_log<<<1, 32>>>(d, c, iters); rep("double log (+2 adds) dependent, 1 warp", iters * 4.0);
    for (int threads : { 32, 128, 256, 512, 1024 }) {
        thr_dfma<<<1, threads>>>(d, c, iters);
        cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("DFMA throughput, 1 CTA of %4d threads x 8 chains: %6.2f DFMA lanes per clock per SM\n", threads,
               (double)threads * iters * 32.0 / (double)h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
