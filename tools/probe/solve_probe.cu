// Isolated timing of the small-system triangular solves and factorisation used by nprior.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NPS_A = 64, NB = 256;
struct Small { double L[NPS_A][NPS_A + 1]; double tmu[NPS_A]; double wt[NPS_A]; };

__global__ void __launch_bounds__(NB) probe(int A, long long* cyc, double* out)
{
    __shared__ Small S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < A * A; e += NB) { const int r = e / A, c = e % A; S.L[r][c] = (r == c) ? 4.0 + r : 0.01 * ((r * 7 + c * 3) % 11); }
    if (tid < A) { S.tmu[tid] = 1.0 + tid; S.wt[tid] = 0.5 * tid; }
    __syncthreads();
    long long t0 = clock64();
    for (int k = 0; k < A; ++k) {
        const double d = S.L[k][k];
        const double dinv = 1.0 / d;
        for (int r = k + 1 + warp; r < A; r += NB / 32) {
            const double lrk = S.L[r][k] * dinv;
            for (int cc = k + 1 + lane; cc <= r; cc += 32) S.L[r][cc] = fma(-lrk, S.L[cc][k], S.L[r][cc]);
        }
        __syncthreads();
    }
    for (int e = tid; e < A * A; e += NB) { const int r = e / A, cc = e % A; if (cc < r) S.L[r][cc] *= rsqrt(S.L[cc][cc]); }
    __syncthreads();
    if (tid < A) S.L[tid][tid] = sqrt(S.L[tid][tid]);
    __syncthreads();
    long long t1 = clock64();
    long long t2 = t1, t3 = t1;
    if (warp == 0) {
        const unsigned full = 0xffffffffu;
        const int r0 = lane, r1 = lane + 32;
        const double id0 = r0 < A ? 1.0 / S.L[r0][r0] : 0.0, id1 = r1 < A ? 1.0 / S.L[r1][r1] : 0.0;
        double t0v = r0 < A ? S.tmu[r0] : 0.0, t1v = r1 < A ? S.tmu[r1] : 0.0;
        for (int k = 0; k < A; ++k) {
            const double xk = __shfl_sync(full, k < 32 ? t0v * id0 : t1v * id1, k & 31);
            if (r0 == k) t0v = xk;
            if (r1 == k) t1v = xk;
            if (r0 > k && r0 < A) t0v = fma(-S.L[r0][k], xk, t0v);
            if (r1 > k && r1 < A) t1v = fma(-S.L[r1][k], xk, t1v);
        }
        t2 = clock64();
        double z0 = r0 < A ? S.wt[r0] : 0.0, z1 = r1 < A ? S.wt[r1] : 0.0;
        for (int k = A - 1; k >= 0; --k) {
            const double mk = __shfl_sync(full, k < 32 ? t0v * id0 : t1v * id1, k & 31);
            const double sk = __shfl_sync(full, k < 32 ? z0 * id0 : z1 * id1, k & 31);
            if (r0 == k) { t0v = mk; z0 = sk; }
            if (r1 == k) { t1v = mk; z1 = sk; }
            if (r0 < k) { const double l = S.L[k][r0]; t0v = fma(-l, mk, t0v); z0 = fma(-l, sk, z0); }
            if (r1 < k) { const double l = S.L[k][r1]; t1v = fma(-l, mk, t1v); z1 = fma(-l, sk, z1); }
        }
        t3 = clock64();
        if (r0 < A) { S.tmu[r0] = t0v; S.wt[r0] = z0; }
        if (r1 < A) { S.tmu[r1] = t1v; S.wt[r1] = z1; }
    }
    __syncthreads();
    if (tid < A) out[tid] = S.tmu[tid] + S.wt[tid];
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}

int main()
{
    long long* c; double* o; cudaMalloc(&c, 24); cudaMalloc(&o, 8 * 64);
    for (int A : { 30, 30, 60 }) {
        probe<<<1, NB>>>(A, c, o);
        long long h[3]; cudaDeviceSynchronize(); cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
        printf("A=%d: factor %lld cycles (%.0f per column), forward %lld (%.0f per step), backward %lld (%.0f per step)\n", A,
               h[0], (double)h[0] / A, h[1], (double)h[1] / A, h[2], (double)h[2] / A);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
