/*
 * gibbs_oracle.c -- CPU restatement of the reference's single-site
 * spike-and-slab Gibbs sampler.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing in pyblip_b200/
 * links, imports or executes it.
 *
 * What is restated (all paths relative to /root/reference):
 *   orc_sample_spikeslab   pyblip/linear/_linear.pyx:34-234   (_sample_spikeslab,
 *                          linear and probit branches)
 *   update_hparams         pyblip/cython_utils/_update_hparams.pyx:18-92
 *   orc_truncnorm          pyblip/cython_utils/_truncnorm.pyx:42-107
 *
 * Pinning status: PINNED.  oracle/pin_against_reference.py drives the
 * reference's own compiled Cython kernel (oracle/_ref) and this restatement
 * with one injected random stream and requires bit-equal indicators and
 * <=1e-11 relative agreement of betas / hyper-parameters; its outputs are
 * committed under tests/golden/ and re-checked by tests/test_oracle.py.
 *
 * Random streams.  The reference mixes three RNGs (libc rand(), NumPy's
 * global MT19937, scipy frozen distributions) that cannot be reproduced on a
 * device.  Parity therefore runs through the "stream contract" shared by this
 * file, the pin script (which monkey-patches the reference's RNG entry
 * points) and the CUDA kernels:
 *   - main stream (permutation, spike/slab uniform, slab normal, hyper
 *     variates): either INJECTED arrays or, if the array pointer is NULL,
 *     Philox4x32-10 keyed by (seed, chain) with the indices below as counter;
 *   - probit truncated-normal attempts: always Philox-keyed by
 *     (seed, chain, redraw event, observation, attempt).
 * The contract is documented in DESIGN.md section "Stream contract".
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off keeps the arithmetic un-fused like the reference's
 * Cython-generated C (plain x86-64, no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011, Random123 constants)             */
/* ------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                       uint32_t c2, uint32_t c3, uint32_t out[4])
{
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream kinds (counter word 3, high byte) */
enum {
    KIND_SPIKE_U = 0,  /* c0 = position in permutation, c1 = sweep        */
    KIND_SLAB_Z = 1,   /* c0 = position in permutation, c1 = sweep        */
    KIND_PERM = 2,     /* c0 = round-key block (0,1),   c1 = sweep        */
    KIND_HYPER = 3,    /* c0 = slot, c1 = sweep, low 24 bits = attempt    */
    KIND_TRUNCNORM = 4 /* c0 = observation, c1 = event, low 24 = attempt  */
};
/* hyper slots */
enum { SLOT_INVGAMMA = 0, SLOT_TAU_GAMMA = 1, SLOT_BETA_A = 10, SLOT_BETA_B = 11 };

typedef struct { uint32_t k0, k1, chain; } orc_key;

static inline double u52(uint32_t a, uint32_t b)
{   /* 52-bit uniform strictly inside (0,1) */
    double v = (double)(a >> 6) * 67108864.0 + (double)(b >> 6);
    return (v + 0.5) * (1.0 / 4503599627370496.0);
}

static inline double box_muller(const uint32_t w[4])
{
    double u1 = u52(w[0], w[1]);
    double u2 = u52(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

double orc_uniform(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0,
                   uint32_t c1, uint32_t attempt)
{
    uint32_t w[4];
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, chain,
                      (kind << 24) | (attempt & 0xFFFFFFu), w);
    return u52(w[0], w[1]);
}

double orc_normal(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0,
                  uint32_t c1, uint32_t attempt)
{
    uint32_t w[4];
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, chain,
                      (kind << 24) | (attempt & 0xFFFFFFu), w);
    return box_muller(w);
}

/* Marsaglia-Tsang Gamma(shape, 1) on the keyed hyper sub-stream `slot`. */
double orc_gamma(uint64_t seed, uint32_t chain, uint32_t slot, uint32_t sweep,
                 double shape)
{
    double boost = 1.0;
    if (shape < 1.0) {
        double ub = orc_uniform(seed, chain, KIND_HYPER, slot, sweep, 0xFFFFFFu);
        boost = pow(ub, 1.0 / shape);
        shape += 1.0;
    }
    double d = shape - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 4000000u; ++t) {
        double x = orc_normal(seed, chain, KIND_HYPER, slot, sweep, 2 * t);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = orc_uniform(seed, chain, KIND_HYPER, slot, sweep, 2 * t + 1);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d; /* unreachable in practice */
}

double orc_beta(uint64_t seed, uint32_t chain, uint32_t proposal, uint32_t sweep,
                double a, double b)
{
    double ga = orc_gamma(seed, chain, SLOT_BETA_A + 2 * proposal, sweep, a);
    double gb = orc_gamma(seed, chain, SLOT_BETA_B + 2 * proposal, sweep, b);
    return ga / (ga + gb);
}

/* Visiting order of (chain, sweep) under the keyed stream: a 6-round alternating
 * Feistel permutation of Z_a x Z_b (a = ceil(sqrt(p)), b = ceil(p/a)) with cycle
 * walking into [0, p); round keys are two Philox blocks (DESIGN.md "Stream
 * contract", kind PERM).  The reference uses np.random.shuffle (_linear.pyx:132);
 * parity runs inject the reference's permutations instead of using this one. */
static uint32_t perm_round(uint32_t v, uint32_t key, uint32_t radix)
{
    v = (v + key) * 0x9E3779B1u;
    v ^= v >> 15; v *= 0x85EBCA77u;
    v ^= v >> 13; v *= 0xC2B2AE3Du;
    v ^= v >> 16;
    return (uint32_t)(((uint64_t)v * radix) >> 32);
}

void orc_permutation(uint64_t seed, uint32_t chain, uint32_t sweep, int p,
                     int32_t *perm)
{
    uint32_t a = 1;
    while ((uint64_t)a * a < (uint64_t)p) ++a;
    const uint32_t b = ((uint32_t)p + a - 1) / a;
    uint32_t x4[4], y4[4], rk[6];
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 0u, sweep, chain,
                      ((uint32_t)KIND_PERM << 24), x4);
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 1u, sweep, chain,
                      ((uint32_t)KIND_PERM << 24), y4);
    rk[0] = x4[0]; rk[1] = x4[1]; rk[2] = x4[2]; rk[3] = x4[3]; rk[4] = y4[0]; rk[5] = y4[1];
    for (int pos = 0; pos < p; ++pos) {
        uint32_t x = (uint32_t)pos;
        do {
            uint32_t L = x / b, R = x - L * b;
            for (int r = 0; r < 6; r += 2) {
                L += perm_round(R, rk[r], a);     if (L >= a) L -= a;
                R += perm_round(L, rk[r + 1], b); if (R >= b) R -= b;
            }
            x = L * b + R;
        } while (x >= (uint32_t)p);
        perm[pos] = (int32_t)x;
    }
}

/* ------------------------------------------------------------------ */
/* Truncated normal: _truncnorm.pyx:42-107                             */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t seed; uint32_t chain, obs, event, attempt;
    /* injected variates (value-level pin against the reference, oracle/pin_truncnorm.py):
     * when inj_u is set, uniforms are read from inj_u[iu++] in the order the reference's
     * rand()/RAND_MAX calls consume them and normals from inj_z[iz++] in the order of its
     * np.random.randn() calls; otherwise the keyed Philox sub-stream is used. */
    const double *inj_u, *inj_z; int64_t iu, iz, nu, nz; int overflow;
} tn_stream;

static double tn_next_pair(tn_stream *s, double *second)
{   /* the two uniforms of one attempt of the exponential-proposal branch
     * (_truncnorm.pyx:50-53: first the proposal, then the accept test) */
    if (s->inj_u) {
        if (s->iu + 2 > s->nu) { s->overflow = 1; *second = 0.0; return 0.5; }   /* forces acceptance */
        double first = s->inj_u[s->iu];
        *second = s->inj_u[s->iu + 1];
        s->iu += 2;
        return first;
    }
    uint32_t w[4];
    orc_philox4x32_10((uint32_t)s->seed, (uint32_t)(s->seed >> 32), s->obs,
                      s->event, s->chain,
                      ((uint32_t)KIND_TRUNCNORM << 24) | (s->attempt & 0xFFFFFFu), w);
    s->attempt++;
    *second = u52(w[2], w[3]);
    return u52(w[0], w[1]);
}

static double tn_next_normal(tn_stream *s)
{
    if (s->inj_z) {
        if (s->iz + 1 > s->nz) { s->overflow = 1; return 1e300; }                /* forces acceptance */
        return s->inj_z[s->iz++];
    }
    uint32_t w[4];
    orc_philox4x32_10((uint32_t)s->seed, (uint32_t)(s->seed >> 32), s->obs,
                      s->event, s->chain,
                      ((uint32_t)KIND_TRUNCNORM << 24) | (s->attempt & 0xFFFFFFu), w);
    s->attempt++;
    return box_muller(w);
}

/* Z ~ N(0,1) | Z >= a     (_truncnorm.pyx:42-86) */
static double truncnorm_std(double a, tn_stream *s)
{
    if (a >= 0.150) {                       /* _expo_tn_sampler :42-55 */
        for (;;) {
            double u2;
            double y = tn_next_pair(s, &u2);
            y = -1 * log(y) / a;
            double rho = exp(-0.5 * y * y);
            if (u2 <= rho) return y + a;
        }
    } else {                                /* _norm_tn_sampler :57-69 */
        for (;;) {
            double z = tn_next_normal(s);
            if (a >= 0) z = fabs(z);
            if (z >= a) return z;
        }
    }
}

/* sample_truncnorm(mean, var, b, lower_interval)   (_truncnorm.pyx:88-107): the one
 * arithmetic path behind both the keyed and the injected entry point */
static double truncnorm_from(tn_stream *s, double mean, double var, double b, int lower_interval)
{
    double scale = sqrt(var);
    double a = (b - mean) / scale;
    double z;
    if (lower_interval == 0) z = truncnorm_std(a, s);
    else z = -1 * truncnorm_std(-1 * a, s);
    return mean + scale * z;
}

double orc_truncnorm(uint64_t seed, uint32_t chain, uint32_t event, uint32_t obs,
                     double mean, double var, double b, int lower_interval)
{
    tn_stream s = { seed, chain, obs, event, 0, NULL, NULL, 0, 0, 0, 0, 0 };
    return truncnorm_from(&s, mean, var, b, lower_interval);
}

/* `count` consecutive sample_truncnorm calls fed from recorded variate sequences: u[] in the
 * order of the reference's rand()/RAND_MAX calls, z[] in the order of its np.random.randn()
 * calls.  used[0], used[1] return how many of each were consumed; u_at[i], z_at[i] (optional)
 * the offsets at which call i started.  Returns -1 if a sequence runs out. */
int orc_truncnorm_injected(int64_t count, const double *mean, const double *var, const double *b,
                           const int32_t *lower_interval, const double *u, int64_t nu,
                           const double *z, int64_t nz, double *out, int64_t *used,
                           int64_t *u_at, int64_t *z_at)
{
    static const double dummy = 0.5;
    tn_stream s = { 0, 0, 0, 0, 0, u ? u : &dummy, z ? z : &dummy, 0, 0, u ? nu : 0, z ? nz : 0, 0 };
    for (int64_t i = 0; i < count; ++i) {
        if (u_at) u_at[i] = s.iu;
        if (z_at) z_at[i] = s.iz;
        out[i] = truncnorm_from(&s, mean[i], var[i], b[i], lower_interval[i]);
        if (s.overflow) return -1;
    }
    used[0] = s.iu; used[1] = s.iz;
    return 0;
}

/* ------------------------------------------------------------------ */
/* The sampler: _linear.pyx:34-234 + _update_hparams.pyx:18-92         */
/* ------------------------------------------------------------------ */
typedef struct {
    double tau2;  int32_t update_tau2;   double tau2_a0, tau2_b0;
    double sigma2; int32_t update_sigma2; double sigma2_a0, sigma2_b0;
    double p0;    int32_t update_p0;     double min_p0, p0_a0, p0_b0;
} orc_hparams;

typedef struct {
    /* any pointer may be NULL -> keyed Philox stream is used for that slot */
    const int32_t *perm;     /* [N][p]      visiting order of sweep i           */
    const double *u;         /* [N][p]      spike/slab uniform, by position     */
    const double *z;         /* [N][p]      slab normal, by position            */
    const double *invgamma;  /* [N]         InvGamma(n/2+sigma2_a0) variates    */
    const double *p0_draw;   /* [N][nprop]  Beta variates (nprop 1 or 100)      */
    const double *gamma_draw;/* [N]         Gamma(tau2_a0+k/2) variates         */
    int32_t nprop;
} orc_streams;

/* _update_hparams (pyblip/cython_utils/_update_hparams.pyx:18-92) for sweep i:
 * beta_i = row i of betas, r = current residual.  Shared by the single-site and
 * the block sampler (oracle/multi_oracle.c). */
void orc_update_hparams(int i, int n, int p, const double *beta_i, const double *r,
                        const orc_hparams *hp, uint64_t seed, uint32_t chain,
                        const orc_streams *st, double *p0s, double *tau2s, double *sigma2s)
{
    const int max_nprop = 100;
    const double sigma_a = n / 2.0 + hp->sigma2_a0;  /* _linear.pyx:108 */
    int num_active = 0;                              /* :52-54 */
    for (int j = 0; j < p; ++j) if (beta_i[j] != 0) num_active++;
    if (hp->update_p0 == 1) {                        /* :57-72 */
        double a = hp->p0_a0 + p - num_active, b = hp->p0_b0 + num_active;
        if (hp->min_p0 == 0) {
            p0s[i] = (st && st->p0_draw) ? st->p0_draw[(size_t)i * st->nprop]
                                         : orc_beta(seed, chain, 0, i, a, b);
        } else {
            p0s[i] = hp->min_p0;
            for (int jj = 0; jj < max_nprop; ++jj) {
                double prop = (st && st->p0_draw)
                    ? st->p0_draw[(size_t)i * st->nprop + jj]
                    : orc_beta(seed, chain, jj, i, a, b);
                if (prop > hp->min_p0) { p0s[i] = prop; break; }
            }
        }
    }
    if (hp->update_sigma2 == 1) {                    /* :75-81 */
        double ss = 0;
        for (int it = 0; it < n; ++it) ss += r[it] * r[it];
        double r2 = sqrt(ss);                        /* dnrm2 ... */
        r2 = r2 * r2;                                /* ... squared, :77-78 */
        double sigma_b = r2 / 2.0 + hp->sigma2_b0;
        double ig = (st && st->invgamma) ? st->invgamma[i]
            : 1.0 / orc_gamma(seed, chain, SLOT_INVGAMMA, i, sigma_a);
        sigma2s[i] = sigma_b * ig;
    }
    if (hp->update_tau2) {                           /* :84-91 */
        double sample_var = 0;
        for (int j = 0; j < p; ++j)
            if (beta_i[j] != 0) sample_var += beta_i[j] * beta_i[j];
        double g = (st && st->gamma_draw) ? st->gamma_draw[i]
            : orc_gamma(seed, chain, SLOT_TAU_GAMMA, i,
                        hp->tau2_a0 + (double)num_active / 2.0);
        tau2s[i] = (hp->tau2_b0 + sample_var / 2.0) / g;
    }
}

int orc_sample_spikeslab(
    int N, int n, int p,
    const double *X,          /* [n][p] row-major, like the reference          */
    const double *y_in,       /* [n]  (ignored if probit)                       */
    const int64_t *zlab,      /* [n]  0/1 labels (probit only)                  */
    int probit,
    const orc_hparams *hp,
    uint64_t seed, uint32_t chain,
    const orc_streams *st,
    double *betas,            /* [N][p] */
    double *p0s, double *tau2s, double *sigma2s, /* [N] */
    double *y_latent,         /* [N][n] or NULL */
    int64_t *n_events_out)    /* probit: number of latent redraw events */
{
    double *XT = (double *)malloc(sizeof(double) * (size_t)n * p);
    double *Xl2 = (double *)calloc(p, sizeof(double));
    double *logdets = (double *)calloc(p, sizeof(double));
    double *post_vars = (double *)calloc(p, sizeof(double));
    double *mu = (double *)calloc(n, sizeof(double));
    double *r = (double *)calloc(n, sizeof(double));
    double *y = (double *)calloc(n, sizeof(double));
    int32_t *inds = (int32_t *)malloc(sizeof(int32_t) * p);
    if (!XT || !Xl2 || !logdets || !post_vars || !mu || !r || !y || !inds) return 1;

    for (int i = 0; i < n; ++i)                      /* :73-74 */
        for (int j = 0; j < p; ++j) {
            double x = X[(size_t)i * p + j];
            XT[(size_t)j * n + i] = x;
            Xl2[j] += x * x;
        }
    memset(betas, 0, sizeof(double) * (size_t)N * p);
    if (y_latent) memset(y_latent, 0, sizeof(double) * (size_t)N * n);
    if (!probit) memcpy(y, y_in, sizeof(double) * n);

    double logodds = log(hp->p0) - log(1 - hp->p0);  /* :76 */
    uint32_t event = 0;

    sigma2s[0] = hp->sigma2; tau2s[0] = hp->tau2; p0s[0] = hp->p0;  /* :112-114 */

    if (probit == 1) {                               /* :117-122 */
        for (int it = 0; it < n; ++it) {
            double v = orc_truncnorm(seed, chain, event, (uint32_t)it, mu[it],
                                     hp->sigma2, 0.0, (int)zlab[it]);
            if (y_latent) y_latent[it] = v;
            y[it] = v;
        }
        event++;
    }
    for (int it = 0; it < n; ++it) r[it] = y[it] - mu[it];  /* :123-124 */

    for (int i = 0; i < N; ++i) {
        double *beta_i = betas + (size_t)i * p;
        for (int j = 0; j < p; ++j) {                /* :128-130 */
            logdets[j] = log(1.0 + tau2s[i] * Xl2[j] / sigma2s[i]) / 2.0;
            post_vars[j] = 1.0 / (1.0 / tau2s[i] + Xl2[j] / sigma2s[i]);
        }
        if (st && st->perm) memcpy(inds, st->perm + (size_t)i * p, sizeof(int32_t) * p);
        else orc_permutation(seed, chain, (uint32_t)i, p, inds);   /* :132 */

        for (int pos = 0; pos < p; ++pos) {          /* :133 */
            int j = inds[pos];
            const double *Xj = XT + (size_t)j * n;
            double old_betaj = beta_i[j];
            if (old_betaj != 0)                      /* :136-137 daxpy */
                for (int it = 0; it < n; ++it) r[it] += old_betaj * Xj[it];
            double XjTr = 0;                         /* :139-145 ddot */
            for (int it = 0; it < n; ++it) XjTr += r[it] * Xj[it];
            double log_ratio_num = tau2s[i] / sigma2s[i] * XjTr * XjTr;  /* :146 */
            double log_ratio_denom = 2 * (sigma2s[i] + tau2s[i] * Xl2[j]);
            double ratio = exp(logodds - log_ratio_num / log_ratio_denom + logdets[j]);
            double kappa = ratio / (1.0 + ratio);
            double u = (st && st->u) ? st->u[(size_t)i * p + pos]
                                     : orc_uniform(seed, chain, KIND_SPIKE_U, pos, i, 0);
            if (u <= kappa) {                        /* :152-153 (NaN -> slab) */
                beta_i[j] = 0;
            } else {                                 /* :155-158 */
                double post_mean = post_vars[j] * XjTr / sigma2s[i];
                double zn = (st && st->z) ? st->z[(size_t)i * p + pos]
                                          : orc_normal(seed, chain, KIND_SLAB_Z, pos, i, 0);
                beta_i[j] = sqrt(post_vars[j]) * zn + post_mean;
            }
            if (probit != 1) {                       /* :160-170 */
                if (beta_i[j] != 0) {
                    double neg = -1 * beta_i[j];
                    for (int it = 0; it < n; ++it) r[it] += neg * Xj[it];
                }
            } else {                                 /* :172-190 */
                double delta = beta_i[j] - old_betaj;
                if (delta != 0) {
                    for (int it = 0; it < n; ++it) mu[it] += delta * Xj[it];
                    for (int it = 0; it < n; ++it) {
                        double v = orc_truncnorm(seed, chain, event, (uint32_t)it,
                                                 mu[it], hp->sigma2, 0.0, (int)zlab[it]);
                        if (y_latent) y_latent[(size_t)i * n + it] = v;
                        y[it] = v;
                        r[it] = y[it] - mu[it];
                    }
                    event++;
                }
            }
        }

        orc_update_hparams(i, n, p, beta_i, r, hp, seed, chain, st, p0s, tau2s, sigma2s);
        logodds = log(p0s[i]) - log(1 - p0s[i]);     /* :222 */
        if (i != N - 1) {                            /* :225-229 */
            memcpy(betas + (size_t)(i + 1) * p, beta_i, sizeof(double) * p);
            p0s[i + 1] = p0s[i];
            sigma2s[i + 1] = sigma2s[i];
            tau2s[i + 1] = tau2s[i];
        }
    }
    if (n_events_out) *n_events_out = (int64_t)event;
    free(XT); free(Xl2); free(logdets); free(post_vars);
    free(mu); free(r); free(y); free(inds);
    return 0;
}
