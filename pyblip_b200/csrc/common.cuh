// common.cuh -- shared device/host helpers for libblipgpu (sm_100a only).
//
// Random-stream contract (DESIGN.md "Stream contract"): every variate is a pure
// function of (seed, chain, kind, index0, index1, attempt) through
// Philox4x32-10, so results do not depend on launch geometry, chunking or the
// number of GPUs.  The CPU oracle implements the same contract independently
// (oracle/gibbs_oracle.c); the device code below is written from the contract,
// not from the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BLIP_OK 0
#define BLIP_ERR_CUDA 1
#define BLIP_ERR_ARG 2
#define BLIP_ERR_NOMEM 3
#define BLIP_ERR_NUMERIC 4   // maps to RuntimeError/ValueError on the Python side
#define BLIP_ERR_UNSUPPORTED 5

void blip_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            blip_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                           cudaGetErrorString(_e));                            \
            return BLIP_ERR_CUDA;                                              \
        }                                                                      \
    } while (0)

#define BLIP_TRY(expr)                                                         \
    do {                                                                       \
        int _rc = (expr);                                                      \
        if (_rc != BLIP_OK) return _rc;                                        \
    } while (0)

// ---------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------
enum : uint32_t {
    KIND_SPIKE_U = 0,   // c0 = position in permutation, c1 = sweep
    KIND_SLAB_Z = 1,    // c0 = position in permutation, c1 = sweep
    KIND_PERM = 2,      // c0 = round-key block (0, 1),  c1 = sweep: keys of the visiting-order permutation
    KIND_HYPER = 3,     // c0 = slot, c1 = sweep, low 24 bits of c3 = attempt
    KIND_TRUNCNORM = 4, // c0 = observation, c1 = redraw event, low 24 = attempt
    KIND_NPRIOR = 5,    // neuronized prior sub-streams (see nprior.cu)
    KIND_BLOCK = 6      // block sampler sub-streams (see gibbs_block.cu)
};
enum : uint32_t { SLOT_INVGAMMA = 0, SLOT_TAU_GAMMA = 1, SLOT_BETA_A = 10, SLOT_BETA_B = 11 };

struct RngKey { uint32_t k0, k1, chain; };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0,
                                                        uint32_t c1, uint32_t c2, uint32_t c3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Visiting order of a sweep: a keyed pseudo-random permutation of [0, p), evaluated point-wise
// (any thread can compute the coordinate visited at any position, no sequential shuffle):
// a 6-round alternating Feistel network on Z_a x Z_b (a = ceil(sqrt(p)), b = ceil(p / a), so
// that a*b exceeds p by less than a) with cycle walking back into [0, p).  Round keys are
// Philox words of (seed, chain, sweep); a round adds a keyed hash of one half to the other
// half modulo its radix (multiply-shift range reduction, no division).  The reference
// shuffles with NumPy's Fisher-Yates (pyblip/linear/_linear.pyx:132); any permutation law
// gives a valid sampler, and parity runs inject the reference's own permutations.
struct PermKey { uint32_t rk[6]; uint32_t a, b; float inv_b; };

__host__ __device__ __forceinline__ PermKey perm_key(uint32_t k0, uint32_t k1, uint32_t chain,
                                                     uint32_t sweep, uint32_t p)
{
    PermKey K;
    uint32_t a = (uint32_t)sqrt((double)p);          // a = ceil(sqrt(p)); p <= 2^31: a <= 46341
    while ((uint64_t)a * a < (uint64_t)p) ++a;
    while (a > 1 && (uint64_t)(a - 1) * (a - 1) >= (uint64_t)p) --a;
    K.a = a;
    K.b = (p + a - 1) / a;
    K.inv_b = 1.0f / (float)K.b;
    const uint4 x = philox4x32_10(k0, k1, 0u, sweep, chain, 2u << 24);
    const uint4 y = philox4x32_10(k0, k1, 1u, sweep, chain, 2u << 24);
    K.rk[0] = x.x; K.rk[1] = x.y; K.rk[2] = x.z; K.rk[3] = x.w; K.rk[4] = y.x; K.rk[5] = y.y;
    return K;
}

__host__ __device__ __forceinline__ uint32_t perm_round(uint32_t v, uint32_t key, uint32_t radix)
{
    v = (v + key) * 0x9E3779B1u;
    v ^= v >> 15; v *= 0x85EBCA77u;
    v ^= v >> 13; v *= 0xC2B2AE3Du;
    v ^= v >> 16;
    return (uint32_t)(((uint64_t)v * radix) >> 32);   // uniform on [0, radix)
}

__host__ __device__ __forceinline__ uint32_t perm_at(const PermKey& K, uint32_t pos, uint32_t p)
{
    uint32_t x = pos;
    do {
        // L = x / b, R = x % b without an integer division: x < a*b < 2^31.5 but a float quotient is
        // only good to a few units, so it is corrected (at most twice each way)
        uint32_t L = (uint32_t)((float)x * K.inv_b);
        if (L >= K.a) L = K.a - 1;
        while ((uint64_t)L * K.b > x) --L;
        while ((uint64_t)(L + 1) * K.b <= x) ++L;
        uint32_t R = x - L * K.b;                     // L in [0,a), R in [0,b)
#pragma unroll
        for (int r = 0; r < 6; r += 2) {
            L += perm_round(R, K.rk[r], K.a);     if (L >= K.a) L -= K.a;
            R += perm_round(L, K.rk[r + 1], K.b); if (R >= K.b) R -= K.b;
        }
        x = L * K.b + R;
    } while (x >= p);
    return x;
}

__host__ __device__ __forceinline__ double u52(uint32_t a, uint32_t b)
{   // uniform strictly inside (0,1) with 52 random bits
    double v = (double)(a >> 6) * 67108864.0 + (double)(b >> 6);
    return (v + 0.5) * (1.0 / 4503599627370496.0);
}

__device__ __forceinline__ double box_muller(uint4 w)
{
    double u1 = u52(w.x, w.y);
    double u2 = u52(w.z, w.w);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

__device__ __forceinline__ uint4 rng_words(const RngKey& key, uint32_t kind, uint32_t c0,
                                           uint32_t c1, uint32_t attempt)
{
    return philox4x32_10(key.k0, key.k1, c0, c1, key.chain, (kind << 24) | (attempt & 0xFFFFFFu));
}
__device__ __forceinline__ double rng_uniform(const RngKey& key, uint32_t kind, uint32_t c0,
                                              uint32_t c1, uint32_t attempt)
{
    uint4 w = rng_words(key, kind, c0, c1, attempt);
    return u52(w.x, w.y);
}
__device__ __forceinline__ double rng_normal(const RngKey& key, uint32_t kind, uint32_t c0,
                                             uint32_t c1, uint32_t attempt)
{
    return box_muller(rng_words(key, kind, c0, c1, attempt));
}

// Marsaglia-Tsang Gamma(shape,1) on hyper sub-stream `slot` of sweep `sweep`.
__device__ inline double rng_gamma(const RngKey& key, uint32_t slot, uint32_t sweep, double shape,
                                   uint32_t kind = KIND_HYPER)
{
    double boost = 1.0;
    if (shape < 1.0) {
        double ub = rng_uniform(key, kind, slot, sweep, 0xFFFFFFu);
        boost = pow(ub, 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 100000u; ++t) {
        double x = rng_normal(key, kind, slot, sweep, 2 * t);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = rng_uniform(key, kind, slot, sweep, 2 * t + 1);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
}

__device__ inline double rng_beta(const RngKey& key, uint32_t proposal, uint32_t sweep, double a,
                                  double b, uint32_t kind = KIND_HYPER)
{
    double ga = rng_gamma(key, SLOT_BETA_A + 2 * proposal, sweep, a, kind);
    double gb = rng_gamma(key, SLOT_BETA_B + 2 * proposal, sweep, b, kind);
    return ga / (ga + gb);
}

// Z ~ N(0,1) | Z >= a.  Exponential-proposal rejection for a >= 0.15, normal
// rejection otherwise: the algorithm of the reference's truncated-normal
// primitive (/root/reference/pyblip/cython_utils/_truncnorm.pyx:42-86).  The arithmetic is
// written once, over a variate source:
//   TnKeyed     attempts drawn from the keyed sub-stream (kind, c0, c1) -- the product path;
//               `first_normal`, when given, is box_muller(rng_words(key, kind, c0, c1, 0)) computed
//               ahead of time by the caller (the first attempt of the normal-rejection branch does
//               not depend on a);
//   TnInjected  uniforms / normals read from recorded sequences in consumption order -- the
//               value-level pin against the reference's own rand() / randn() draws
//               (blipgpu_debug_truncnorm, oracle/pin_truncnorm.py).
// Attempts are capped so that a NaN argument (upstream bug) cannot hang the GPU.
struct TnKeyed {
    RngKey key; uint32_t kind, c0, c1; const double* first_normal; uint32_t attempt;
    static constexpr uint32_t kMaxAttempts = 1u << 14;
    __device__ __forceinline__ bool pair(double& y, double& u2) {
        if (attempt >= kMaxAttempts) return false;
        const uint4 w = rng_words(key, kind, c0, c1, attempt++);
        y = u52(w.x, w.y); u2 = u52(w.z, w.w);
        return true;
    }
    __device__ __forceinline__ bool normal(double& z) {
        if (attempt >= kMaxAttempts) return false;
        if (attempt == 0 && first_normal) { z = *first_normal; attempt = 1; return true; }
        z = box_muller(rng_words(key, kind, c0, c1, attempt++));
        return true;
    }
};
struct TnInjected {
    const double* u; const double* z; long long iu, iz, nu, nz;
    __device__ __forceinline__ bool pair(double& y, double& u2) {
        if (iu + 2 > nu) return false;
        y = u[iu]; u2 = u[iu + 1]; iu += 2;
        return true;
    }
    __device__ __forceinline__ bool normal(double& zz) {
        if (iz + 1 > nz) return false;
        zz = z[iz++];
        return true;
    }
};

template <class Src>
__device__ __forceinline__ double truncnorm_std_from(Src& src, double a)
{
    if (a >= 0.150) {                        // _expo_tn_sampler :42-55
        double y, u2;
        while (src.pair(y, u2)) {
            y = -1 * log(y) / a;
            double rho = exp(-0.5 * y * y);
            if (u2 <= rho) return y + a;
        }
    } else {                                 // _norm_tn_sampler :57-69
        double z;
        while (src.normal(z)) {
            if (a >= 0) z = fabs(z);
            if (z >= a) return z;
        }
    }
    return nan("");
}

// sample_truncnorm(mean, var, b, lower_interval)  (_truncnorm.pyx:88-107)
template <class Src>
__device__ __forceinline__ double truncnorm_from(Src& src, double mean, double var, double b, int lower_interval)
{
    double scale = sqrt(var);
    double a = (b - mean) / scale;
    double z;
    if (lower_interval == 0) z = truncnorm_std_from(src, a);
    else z = -1 * truncnorm_std_from(src, -1 * a);
    return mean + scale * z;
}

__device__ __forceinline__ double truncnorm_std(const RngKey& key, uint32_t kind, uint32_t c0, uint32_t c1,
                                                double a, const double* first_normal = nullptr)
{
    TnKeyed src{ key, kind, c0, c1, first_normal, 0u };
    return truncnorm_std_from(src, a);
}

__device__ __forceinline__ double truncnorm(const RngKey& key, uint32_t kind, uint32_t c0, uint32_t c1,
                                            double mean, double var, double b, int lower_interval,
                                            const double* first_normal = nullptr)
{
    TnKeyed src{ key, kind, c0, c1, first_normal, 0u };
    return truncnorm_from(src, mean, var, b, lower_interval);
}

// ---------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Barrier policies: the whole CTA, or a named barrier over a warp-specialised sub-group of NT
// threads (warps [0, NT/32) by convention of the callers, which index scratch by threadIdx.x >> 5).
struct SyncCta {
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ int nwarp() { return (blockDim.x + 31) >> 5; }
    static __device__ __forceinline__ int tid() { return threadIdx.x; }
};
// the sub-group is threads [T0, T0 + NT) of the CTA; tid() is the index within it
template <int ID, int NT, int T0 = 0> struct SyncNamed {
    static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory"); }
    static __device__ __forceinline__ int nwarp() { return NT / 32; }
    static __device__ __forceinline__ int tid() { return (int)threadIdx.x - T0; }
};

// Sum over the group SY synchronises (default: the whole block); result returned to every thread.
// `scratch` must hold >= 33 doubles of shared memory.  Contains two barriers.
template <class SY = SyncCta>
__device__ __forceinline__ double block_sum(double v, double* scratch)
{
    const int lane = SY::tid() & 31, warp = SY::tid() >> 5, nwarp = SY::nwarp();
    v = warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    SY::sync();
    if (warp == 0) {
        double t = (lane < nwarp) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    SY::sync();
    return scratch[32];
}
