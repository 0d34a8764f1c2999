/*
 * nprior_oracle.c -- CPU restatement of the reference's neuronized-prior Gibbs sampler.
 * TEST INFRASTRUCTURE ONLY (same rule as gibbs_oracle.c: only tests/, smoke() and
 * bench.py's CPU legs may load this).
 *
 * Restates /root/reference/pyblip/nprior/_nprior.pyx:
 *   orc_nprior_sample         :56-260   _nprior_sample
 *   sample_w                  :266-444  _sample_w (joint draw of W on the active set)
 *   (inline)                  :451-469  _sample_p0, :476-502 _sample_delta,
 *                             :510-537  _sample_sigma2, :543-586 _update_wj,
 *                             :593-678  _update_coordinate_relu
 *   norm_quantile_apprx       :40-50 with rationalapprox (_truncnorm.pyx:22-38)
 *
 * Pinning status: PINNED by oracle/pin_nprior.py (the reference's compiled kernel run with
 * its native NumPy / libc streams, every variate recorded and replayed here; truncated
 * normals keyed on both sides) -> tests/golden/nprior_*.npz.
 *
 * Random streams: as in gibbs_oracle.c.  Extra kinds (counter word 3, high byte):
 *   16 initial alpha_j        c0 = j          17 fresh w row of sweep i   c0 = j, c1 = i
 *   18 initial sigma2 normal                  19 mixture uniform          c0 = pos, c1 = i
 *   20 alpha_j truncated normal attempts      c0 = pos, c1 = i
 *   21 w_j normal             c0 = pos, c1 = i
 *   22 joint-W normals of all p coordinates   c0 = j, c1 = i
 *   23 joint-W normals of the active set      c0 = index in active set, c1 = i
 *   24 group-move normal      c1 = i
 * invgamma / Beta reuse KIND_HYPER slots 0 and 10/11; the permutation is KIND_PERM.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

double orc_uniform(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1, uint32_t attempt);
double orc_normal(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1, uint32_t attempt);
double orc_gamma(uint64_t seed, uint32_t chain, uint32_t slot, uint32_t sweep, double shape);
double orc_beta(uint64_t seed, uint32_t chain, uint32_t proposal, uint32_t sweep, double a, double b);
void orc_permutation(uint64_t seed, uint32_t chain, uint32_t sweep, int p, int32_t *perm);
void orc_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);

enum { NP_INIT_ALPHA = 16, NP_W_FRESH = 17, NP_SIGMA_INIT = 18, NP_U = 19, NP_TN = 20, NP_WZ = 21,
       NP_JOINT_ALL = 22, NP_JOINT_ACT = 23, NP_DELTA = 24 };

/* generic keyed truncated normal: same algorithm as orc_truncnorm (_truncnorm.pyx:42-107),
 * attempts on sub-stream (kind, c0, c1) */
static double u52(uint32_t a, uint32_t b)
{
    double v = (double)(a >> 6) * 67108864.0 + (double)(b >> 6);
    return (v + 0.5) * (1.0 / 4503599627370496.0);
}
static double tn_std(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1, double a)
{
    uint32_t w[4], attempt = 0;
    if (a >= 0.150) {
        for (;;) {
            orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, chain, (kind << 24) | (attempt++ & 0xFFFFFFu), w);
            double y = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
            y = -1 * log(y) / a;
            double rho = exp(-0.5 * y * y);
            if (u2 <= rho) return y + a;
        }
    } else {
        for (;;) {
            orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, chain, (kind << 24) | (attempt++ & 0xFFFFFFu), w);
            double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
            if (a >= 0) z = fabs(z);
            if (z >= a) return z;
        }
    }
}
double orc_truncnorm_keyed(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1,
                           double mean, double var, double b, int lower_interval)
{
    double scale = sqrt(var);
    double a = (b - mean) / scale;
    double z = lower_interval == 0 ? tn_std(seed, chain, kind, c0, c1, a)
                                   : -1 * tn_std(seed, chain, kind, c0, c1, -1 * a);
    return mean + scale * z;
}

static double rationalapprox(double t)            /* _truncnorm.pyx:22-38 */
{
    double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328, d0 = 1.432788, d1 = 0.189269, d2 = 0.001308;
    double frac = (c2 * t + c1) * t + c0;
    frac = frac / (((d2 * t + d1) * t + d0) * t + 1.0);
    return t - frac;
}
double orc_norm_quantile_apprx(double p)          /* _nprior.pyx:40-50 */
{
    if (p < 0.5) return -1 * rationalapprox(sqrt(-2.0 * log(p)));
    return rationalapprox(sqrt(-2.0 * log(1 - p)));
}
static double normcdf(double z) { return 0.5 * erfc(-z * sqrt(0.5)); }   /* :37-38 */

typedef struct {
    /* any pointer may be NULL -> keyed stream */
    const int32_t *perm;        /* [N][p]                                            */
    const double *u;            /* [N][p] mixture uniform, by position               */
    const double *wz;           /* [N][p] normal of the w_j update, by position      */
    const double *alpha_init;   /* [p]    randn for alphas[0] (before scaling)       */
    const double *w_fresh;      /* [N][p] randn for ws_arr (before the 0.1 scale)    */
    const double *sigma_init;   /* [1]                                               */
    const double *joint_all;    /* [N][p] randn(p) inside _sample_w                  */
    const double *joint_act;    /* [N][p] randn(num_active) inside _sample_w, padded */
    const double *delta_z;      /* [N]    group-move normal                          */
    const double *invgamma;     /* [N]                                               */
    const double *p0_draw;      /* [N]    Beta variates                              */
} orc_nprior_streams;

/* _sample_w (:266-444).  Work arrays sized p. */
static int sample_w(int n, int p, const double *XT, const double *y, const double *alpha, double sigma2,
                    double alpha0, double tauw2, uint64_t seed, uint32_t chain, int sweep,
                    const orc_nprior_streams *st, double *w_out)
{
    int num_active = 0;
    double *T = (double *)calloc(p, sizeof(double));
    for (int j = 0; j < p; ++j) {
        double z = (st && st->joint_all) ? st->joint_all[(size_t)sweep * p + j]
                                         : orc_normal(seed, chain, NP_JOINT_ALL, j, sweep, 0);
        w_out[j] = z / sqrt(sigma2 / tauw2);                          /* :288 */
        T[j] = fmax(alpha[j] - alpha0, 0);
        if (T[j] > 0) num_active++;
    }
    if (num_active == 0) { free(T); return 0; }
    int A = num_active;
    double *Phi = (double *)calloc((size_t)A * n, sizeof(double));
    double *PTP = (double *)calloc((size_t)A * A, sizeof(double));
    double *tmu = (double *)calloc(A, sizeof(double));
    double *wt = (double *)calloc(A, sizeof(double));
    int a = 0;
    for (int j = 0; j < p; ++j)
        if (T[j] > 0) {
            for (int i = 0; i < n; ++i) Phi[(size_t)a * n + i] = T[j] * XT[(size_t)j * n + i];
            a++;
        }
    for (int r = 0; r < A; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += Phi[(size_t)r * n + i] * Phi[(size_t)c * n + i];
            PTP[(size_t)r * A + c] = PTP[(size_t)c * A + r] = s;
        }
    for (int r = 0; r < A; ++r) PTP[(size_t)r * A + r] += 1 / tauw2;  /* :336-337 */
    /* Cholesky PTP = L L^T, L lower, in place (lower triangle) */
    for (int k = 0; k < A; ++k) {
        double d = PTP[(size_t)k * A + k];
        for (int m = 0; m < k; ++m) d -= PTP[(size_t)k * A + m] * PTP[(size_t)k * A + m];
        if (!(d > 0)) { free(T); free(Phi); free(PTP); free(tmu); free(wt); return 1; }
        d = sqrt(d);
        PTP[(size_t)k * A + k] = d;
        for (int r = k + 1; r < A; ++r) {
            double s = PTP[(size_t)r * A + k];
            for (int m = 0; m < k; ++m) s -= PTP[(size_t)r * A + m] * PTP[(size_t)k * A + m];
            PTP[(size_t)r * A + k] = s / d;
        }
    }
    /* tmu = Phi y; then (L L^T)^{-1} tmu by two triangular solves (:371-415) */
    for (int r = 0; r < A; ++r) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += Phi[(size_t)r * n + i] * y[i];
        tmu[r] = s;
    }
    for (int r = 0; r < A; ++r) {                /* L v = tmu */
        double s = tmu[r];
        for (int m = 0; m < r; ++m) s -= PTP[(size_t)r * A + m] * tmu[m];
        tmu[r] = s / PTP[(size_t)r * A + r];
    }
    for (int r = A - 1; r >= 0; --r) {           /* L^T m = v */
        double s = tmu[r];
        for (int m = r + 1; m < A; ++m) s -= PTP[(size_t)m * A + r] * tmu[m];
        tmu[r] = s / PTP[(size_t)r * A + r];
    }
    for (int r = 0; r < A; ++r)
        wt[r] = (st && st->joint_act) ? st->joint_act[(size_t)sweep * p + r]
                                      : orc_normal(seed, chain, NP_JOINT_ACT, r, sweep, 0);
    for (int r = A - 1; r >= 0; --r) {           /* L^T s = z  (wt <- L^{-T} wt, :417-432) */
        double s = wt[r];
        for (int m = r + 1; m < A; ++m) s -= PTP[(size_t)m * A + r] * wt[m];
        wt[r] = s / PTP[(size_t)r * A + r];
    }
    a = 0;
    for (int j = 0; j < p; ++j)
        if (T[j] > 0) { w_out[j] = wt[a] + tmu[a]; a++; }
    free(T); free(Phi); free(PTP); free(tmu); free(wt);
    return 0;
}

int orc_nprior_sample(
    int N, int n, int p, const double *X, const double *y,
    double tauw2, double p0_init, double min_p0, int update_p0, double sigma_a0, double sigma_b0,
    int sigma_prior_type, int joint_sample_W, int group_alpha_update,
    uint64_t seed, uint32_t chain, const orc_nprior_streams *st,
    double *alphas, double *ws, double *betas, double *sigma2s, double *alpha0s, double *p0s)
{
    double *XT = (double *)malloc(sizeof(double) * (size_t)n * p);
    double *Xl2 = (double *)calloc(p, sizeof(double));
    double *r = (double *)calloc(n, sizeof(double));
    int32_t *inds = (int32_t *)malloc(sizeof(int32_t) * p);
    if (!XT || !Xl2 || !r || !inds) return 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) {
            double x = X[(size_t)i * p + j];
            XT[(size_t)j * n + i] = x;
            Xl2[j] += x * x;
        }
    double min_alpha0 = orc_norm_quantile_apprx(min_p0);             /* :83 */
    double sigma_a = n / 2.0 + sigma_a0;                             /* :91-93 */
    if (sigma_prior_type != 0) sigma_a += p / 2.0;

    memset(betas, 0, sizeof(double) * (size_t)N * p);
    for (int j = 0; j < p; ++j) {                                    /* :98 (only row 0 survives) */
        double z = (st && st->alpha_init) ? st->alpha_init[j] : orc_normal(seed, chain, NP_INIT_ALPHA, j, 0, 0);
        alphas[j] = 1 / sqrt(2 * log(p)) * z;
    }
    for (int i = 0; i < N; ++i)                                      /* :100 every row is fresh noise */
        for (int j = 0; j < p; ++j) {
            double z = (st && st->w_fresh) ? st->w_fresh[(size_t)i * p + j] : orc_normal(seed, chain, NP_W_FRESH, j, i, 0);
            ws[(size_t)i * p + j] = 0.1 * z;
        }
    {
        double z = (st && st->sigma_init) ? st->sigma_init[0] : orc_normal(seed, chain, NP_SIGMA_INIT, 0, 0, 0);
        sigma2s[0] = z * z;                                          /* :117 */
    }
    p0s[0] = p0_init;
    alpha0s[0] = orc_norm_quantile_apprx(p0_init);

    for (int i = 0; i < N; ++i) {
        double *al = alphas + (size_t)i * p, *w = ws + (size_t)i * p, *b = betas + (size_t)i * p;
        if (joint_sample_W == 1) {
            if (sample_w(n, p, XT, y, al, sigma2s[i], alpha0s[i], tauw2, seed, chain, i, st, w)) return 2;
        }
        for (int j = 0; j < p; ++j) b[j] = w[j] * fmax(al[j] - alpha0s[i], 0);   /* :143-144 */
        for (int it = 0; it < n; ++it) r[it] = y[it];
        for (int j = 0; j < p; ++j)
            if (b[j] != 0) for (int it = 0; it < n; ++it) r[it] -= b[j] * XT[(size_t)j * n + it];

        if (st && st->perm) memcpy(inds, st->perm + (size_t)i * p, sizeof(int32_t) * p);
        else orc_permutation(seed, chain, (uint32_t)i, p, inds);

        for (int pos = 0; pos < p; ++pos) {
            int j = inds[pos];
            const double *Xj = XT + (size_t)j * n;
            double sigma2 = sigma2s[i], alpha0 = alpha0s[i], p0 = p0s[i];
            for (int it = 0; it < n; ++it) r[it] += b[j] * Xj[it];    /* :168 */
            /* ---- _update_coordinate_relu (:593-678) */
            double wj = w[j];
            double r2 = 0;
            for (int it = 0; it < n; ++it) r2 += r[it] * r[it];
            r2 = sqrt(r2); r2 = r2 * r2;
            double r_scale = alpha0 * wj;
            if (r_scale != 0) for (int it = 0; it < n; ++it) r[it] += r_scale * Xj[it];
            double t_denom = wj * wj * Xl2[j] + sigma2;
            double t_sigma2 = sigma2 / t_denom;
            double t_alphaj = 0;
            for (int it = 0; it < n; ++it) t_alphaj += r[it] * Xj[it];
            t_alphaj = wj * t_alphaj / t_denom;
            double rplus2 = 0;
            for (int it = 0; it < n; ++it) rplus2 += r[it] * r[it];
            rplus2 = sqrt(rplus2); rplus2 = rplus2 * rplus2;
            double log_kappa_num = log(p0);
            log_kappa_num -= r2 / (2 * sigma2);
            double log_kappa_denom = log(1 - normcdf((alpha0 - t_alphaj) / sqrt(t_sigma2)));
            log_kappa_denom += log(t_sigma2) / 2 + (t_alphaj * t_alphaj / (2 * t_sigma2) - rplus2 / (2 * sigma2));
            if (r_scale != 0) {
                r_scale = -1 * r_scale;
                for (int it = 0; it < n; ++it) r[it] += r_scale * Xj[it];
            }
            double ratio = exp(log_kappa_num - log_kappa_denom);
            double kappa = ratio / (1.0 + ratio);
            double u = (st && st->u) ? st->u[(size_t)i * p + pos] : orc_uniform(seed, chain, NP_U, pos, i, 0);
            double new_alpha;
            if (u < kappa) new_alpha = orc_truncnorm_keyed(seed, chain, NP_TN, pos, i, 0.0, 1.0, alpha0, 1);
            else new_alpha = orc_truncnorm_keyed(seed, chain, NP_TN, pos, i, t_alphaj, t_sigma2, alpha0, 0);
            al[j] = new_alpha;
            /* ---- _update_wj (:543-586) */
            double Tj = fmax(new_alpha - alpha0, 0);
            double sigma02 = sigma_prior_type == 0 ? 1.0 : sigma2;
            double vj = Xl2[j] * Tj * Tj + (sigma02 / tauw2);
            double mj = 0;
            if (Tj > 0) {
                for (int it = 0; it < n; ++it) mj += r[it] * Xj[it];
                mj = mj * Tj / vj;
            }
            double z = (st && st->wz) ? st->wz[(size_t)i * p + pos] : orc_normal(seed, chain, NP_WZ, pos, i, 0);
            w[j] = sqrt(sigma2 / vj) * z + mj;
            /* ---- :195-198 */
            b[j] = -1 * w[j] * fmax(al[j] - alpha0, 0);
            for (int it = 0; it < n; ++it) r[it] += b[j] * Xj[it];
            b[j] = -1 * b[j];
        }
        /* ---- _sample_sigma2 (:510-537) */
        {
            double r2 = 0;
            for (int it = 0; it < n; ++it) r2 += r[it] * r[it];
            r2 = sqrt(r2); r2 = r2 * r2;
            double bb = r2 / 2 + sigma_b0;
            if (sigma_prior_type != 0) {
                double w2 = 0;
                for (int j = 0; j < p; ++j) w2 += w[j] * w[j];
                w2 = sqrt(w2); w2 = w2 * w2;
                bb = bb + w2 / (2.0 * tauw2);
            }
            double ig = (st && st->invgamma) ? st->invgamma[i] : 1.0 / orc_gamma(seed, chain, 0, i, sigma_a);
            sigma2s[i] = bb * ig;
        }
        if (update_p0 != 0) {
            if (group_alpha_update == 0) {                           /* :219-225, _sample_p0 */
                int k = 0;
                for (int j = 0; j < p; ++j) if (al[j] > alpha0s[i]) k++;
                double v = (st && st->p0_draw) ? st->p0_draw[i] : orc_beta(seed, chain, 0, i, 1 + p - k, 1 + k);
                p0s[i] = fmax(min_p0, v);
                alpha0s[i] = orc_norm_quantile_apprx(p0s[i]);
            } else {                                                 /* :226-236, _sample_delta */
                double min_delta = min_alpha0 - alpha0s[i];
                double mean = 0, pf = p;
                for (int j = 0; j < p; ++j) mean += al[j];
                mean = -1 * (mean + alpha0s[i]) / (pf + 1);
                double z = (st && st->delta_z) ? st->delta_z[i] : orc_normal(seed, chain, NP_DELTA, 0, i, 0);
                double delta = 1 / sqrt(pf + 1) * z + mean;
                delta = fmax(min_delta, delta);
                alpha0s[i] += delta;
                p0s[i] = normcdf(alpha0s[i]);
                for (int j = 0; j < p; ++j) al[j] += delta;
            }
        }
        if (i < N - 1) {                                             /* :239-243 */
            memcpy(alphas + (size_t)(i + 1) * p, al, sizeof(double) * p);
            alpha0s[i + 1] = alpha0s[i];
            p0s[i + 1] = p0s[i];
            sigma2s[i + 1] = sigma2s[i];
        }
    }
    free(XT); free(Xl2); free(r); free(inds);
    return 0;
}
