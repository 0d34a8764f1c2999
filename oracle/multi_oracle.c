/*
 * multi_oracle.c -- CPU restatement of the reference's BLOCK spike-and-slab
 * Gibbs sampler (`bsize > 1`).  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * What is restated (paths relative to /root/reference):
 *   orc_sample_spikeslab_multi   pyblip/linear/_linear_multi.pyx:271-932
 *                                (_sample_spikeslab_multi, linear and probit)
 *   model_logprob / chol / ...   _compute_QA :162-264, _precompute_suff_stats :83-160
 *   weighted_choice              _weighted_choice :55-79
 *
 * Pinning status: PINNED by oracle/pin_multi.py, which drives the reference's
 * compiled kernel (oracle/_ref) and this file with one recorded random stream
 * and requires identical indicators and <=1e-9 relative agreement of betas /
 * hyper-parameters / latents; outputs are committed as tests/golden/multi_*.npz.
 *
 * Reference behaviour that is reproduced on purpose (SURVEY.md section 8, quirks
 * 3 and 4):
 *   - the partition into blocks of `bsize` consecutive (wrapping) indices is
 *     shuffled ONCE: `np.random.shuffle(blocks_arr)` of sweep 0 acts on the
 *     buffer the kernel reads, `blocks_arr = (blocks_arr + 1) % p` then rebinds
 *     the name, so later shuffles/shifts never reach the kernel (:426-427);
 *   - the last block wraps around to index 0 when bsize does not divide p
 *     (:395-402), so those coordinates are visited twice per sweep;
 *   - `_weighted_choice` starts `max_log_prob` at 0.0 and returns 0 when the
 *     cumulative sum never reaches u (:62, :74-77).
 *
 * Random stream (DESIGN.md "Stream contract"): block table, choice uniforms and
 * slab normals are either INJECTED or keyed (kind BLOCK: c0 = block position in
 * the visiting order, c1 = sweep, attempt 0 = uniform, 1+k = k-th normal);
 * hyper variates and probit latents as in gibbs_oracle.c.
 *
 * Small dense algebra is written out (no BLAS/LAPACK): matrices are at most
 * bsize x bsize.  The reference passes C row-major buffers to Fortran LAPACK;
 * all matrices involved are symmetric up to rounding, and the operation ORDER of
 * the reference is kept (Fortran view), see the comments below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { KIND_BLOCK = 6 };

typedef struct {
    double tau2;  int32_t update_tau2;   double tau2_a0, tau2_b0;
    double sigma2; int32_t update_sigma2; double sigma2_a0, sigma2_b0;
    double p0;    int32_t update_p0;     double min_p0, p0_a0, p0_b0;
} orc_hparams;

typedef struct {
    const int32_t *perm; const double *u; const double *z;
    const double *invgamma; const double *p0_draw; const double *gamma_draw;
    int32_t nprop;
} orc_streams;

typedef struct {
    const int32_t *blocks;   /* [nblock][bsize] block table after the sweep-0 shuffle, or NULL */
    const double *u;         /* [N][nblock]        uniform of _weighted_choice               */
    const double *z;         /* [N][nblock][bsize] normals of the selected model, first msize */
    const double *invgamma;  /* [N] */
    const double *p0_draw;   /* [N][nprop] */
    const double *gamma_draw;/* [N] */
    int32_t nprop;
} orc_multi_streams;

double orc_uniform(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1, uint32_t attempt);
double orc_normal(uint64_t seed, uint32_t chain, uint32_t kind, uint32_t c0, uint32_t c1, uint32_t attempt);
double orc_truncnorm(uint64_t seed, uint32_t chain, uint32_t event, uint32_t obs,
                     double mean, double var, double b, int lower_interval);
void orc_permutation(uint64_t seed, uint32_t chain, uint32_t sweep, int p, int32_t *perm);
void orc_update_hparams(int i, int n, int p, const double *beta_i, const double *r,
                        const orc_hparams *hp, uint64_t seed, uint32_t chain,
                        const orc_streams *st, double *p0s, double *tau2s, double *sigma2s);

/* `int nblock = round((p + bsize - 1) / bsize)` (:298): p and bsize are C ints and the
 * function is compiled with cdivision(True), so the division is C integer division and
 * nblock = ceil(p / bsize) (checked against the compiled reference by pin_multi.py: the
 * recorded block table has exactly this many rows). */
int orc_multi_nblock(int p, int bsize)
{
    return (p + bsize - 1) / bsize;
}

/* itertools.combinations order, sizes 1..maxsig (:405-409); returns the number of
 * models written (without the empty model).  masks[k] = bit set of members. */
int orc_multi_models(int bsize, int maxsig, uint32_t *masks, int cap)
{
    int count = 0, idx[32];
    for (int m = 1; m <= maxsig; ++m) {
        for (int t = 0; t < m; ++t) idx[t] = t;
        for (;;) {
            uint32_t mask = 0;
            for (int t = 0; t < m; ++t) mask |= 1u << idx[t];
            if (count < cap && masks) masks[count] = mask;
            count++;
            int t = m - 1;
            while (t >= 0 && idx[t] == bsize - m + t) --t;
            if (t < 0) break;
            idx[t]++;
            for (int s = t + 1; s < m; ++s) idx[s] = idx[s - 1] + 1;
        }
    }
    return count;
}

/* lower Cholesky in place (row-major m x m, leading dimension ld); 0 on success */
static int chol_lower(double *A, int m, int ld)
{
    for (int j = 0; j < m; ++j) {
        double d = A[j * ld + j];
        for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < m; ++i) {
            double s = A[i * ld + j];
            for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = s / d;
        }
    }
    return 0;
}

typedef struct {
    int bsize, m;
    int members[32];
    double A[32 * 32];      /* XATXA of the model (m x m)                         */
    double Linv[32 * 32];   /* inverse of the Cholesky factor of QA = I + c A      */
    double XATr[32];
    double mu[32];          /* Linv * XATr                                         */
    double logdet;
} model_work;

/* _compute_QA (:162-264) for the model `mask` of the current block */
static int compute_QA(model_work *w, uint32_t mask, int bsize, const double *XATXA_cache,
                      const double *XATr_cache, double tau2, double sigma2)
{
    int m = 0;
    for (int jj = 0; jj < bsize; ++jj) if (mask & (1u << jj)) w->members[m++] = jj;
    w->m = m; w->bsize = bsize;
    for (int a = 0; a < m; ++a) {
        w->XATr[a] = XATr_cache[w->members[a]];
        for (int b = 0; b < m; ++b) w->A[a * m + b] = XATXA_cache[w->members[a] * bsize + w->members[b]];
    }
    const double cm = tau2 / sigma2;
    double *Q = w->Linv;
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) Q[a * m + b] = (a == b ? 1.0 : 0.0) + cm * w->A[a * m + b];
    int info = chol_lower(Q, m, m);                       /* dpotrf :229-236 */
    if (info) return info;
    w->logdet = 0.0;
    for (int a = 0; a < m; ++a) w->logdet += 2.0 * log(Q[a * m + a]);   /* :239-241 */
    /* dtrtri :242-249: forward substitution column by column */
    for (int j = 0; j < m; ++j) {
        double dj = 1.0 / Q[j * m + j];
        for (int i = j + 1; i < m; ++i) {
            double s = Q[i * m + j] * dj;
            for (int k = j + 1; k < i; ++k) s += Q[i * m + k] * Q[k * m + j];   /* Q[k][j] already inverted */
            Q[i * m + j] = -s / Q[i * m + i];
        }
        Q[j * m + j] = dj;
    }
    /* zero the strict upper triangle so that Linv is a clean lower-triangular matrix */
    for (int a = 0; a < m; ++a) for (int b = a + 1; b < m; ++b) Q[a * m + b] = 0.0;
    for (int a = 0; a < m; ++a) {                         /* dtrmm :250-263: mu = Linv XATr */
        double s = 0.0;
        for (int k = 0; k <= a; ++k) s += Q[a * m + k] * w->XATr[k];
        w->mu[a] = s;
    }
    return 0;
}

int orc_sample_spikeslab_multi(
    int N, int bsize, int n, int p,
    const double *X, const double *y_in, const int64_t *zlab, int probit,
    const orc_hparams *hp, int max_signals_per_block,
    uint64_t seed, uint32_t chain, const orc_multi_streams *st,
    double *betas, double *p0s, double *tau2s, double *sigma2s, double *y_latent,
    int64_t *n_events_out)
{
    if (bsize < 1 || bsize > 16) return 3;
    const int nblock = orc_multi_nblock(p, bsize);
    if (max_signals_per_block == 0) max_signals_per_block = bsize;     /* :403-404 */
    const int nsub = orc_multi_models(bsize, max_signals_per_block, NULL, 0);
    const int nmodel = nsub + 1;
    uint32_t *masks = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nsub > 0 ? nsub : 1));
    double *XT = (double *)malloc(sizeof(double) * (size_t)n * p);
    double *XTX = (double *)malloc(sizeof(double) * (size_t)p * p);
    double *mu_pred = (double *)calloc(n, sizeof(double));
    double *r = (double *)calloc(n, sizeof(double));
    double *y = (double *)calloc(n, sizeof(double));
    int32_t *blocks = (int32_t *)malloc(sizeof(int32_t) * (size_t)nblock * bsize);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)nblock);
    double *lprobs = (double *)malloc(sizeof(double) * (size_t)nmodel);
    double *probs = (double *)malloc(sizeof(double) * (size_t)nmodel);
    if (!masks || !XT || !XTX || !mu_pred || !r || !y || !blocks || !order || !lprobs || !probs) return 1;
    orc_multi_models(bsize, max_signals_per_block, masks, nsub);

    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) XT[(size_t)j * n + i] = X[(size_t)i * p + j];
    for (int a = 0; a < p; ++a)                                        /* XTX = X'X :379 */
        for (int b = a; b < p; ++b) {
            double s = 0.0;
            for (int it = 0; it < n; ++it) s += XT[(size_t)a * n + it] * XT[(size_t)b * n + it];
            XTX[(size_t)a * p + b] = s; XTX[(size_t)b * p + a] = s;
        }
    memset(betas, 0, sizeof(double) * (size_t)N * p);
    if (y_latent) memset(y_latent, 0, sizeof(double) * (size_t)N * n);
    if (!probit) memcpy(y, y_in, sizeof(double) * n);

    double logp0 = log(hp->p0), log1p0 = log(1 - hp->p0);             /* :339-340 */
    sigma2s[0] = hp->sigma2; tau2s[0] = hp->tau2; p0s[0] = hp->p0;    /* :383-385 */
    uint32_t event = 0;
    if (probit == 1) {                                                 /* :387-392 */
        for (int it = 0; it < n; ++it) {
            double v = orc_truncnorm(seed, chain, event, (uint32_t)it, mu_pred[it], hp->sigma2, 0.0, (int)zlab[it]);
            if (y_latent) y_latent[it] = v;
            y[it] = v;
        }
        event++;
    }
    for (int it = 0; it < n; ++it) r[it] = y[it] - mu_pred[it];       /* :393-394 */

    /* block table (:395-402) in the order fixed by the sweep-0 shuffle */
    if (st && st->blocks) memcpy(blocks, st->blocks, sizeof(int32_t) * (size_t)nblock * bsize);
    else {
        orc_permutation(seed, chain, 0u, nblock, order);
        for (int bi = 0; bi < nblock; ++bi)
            for (int bj = 0; bj < bsize; ++bj) blocks[bi * bsize + bj] = (int32_t)(((int64_t)order[bi] * bsize + bj) % p);
    }

    model_work w;
    double XATXA_cache[16 * 16], XATr_cache[16];
    orc_streams hst;
    memset(&hst, 0, sizeof(hst));
    if (st) { hst.invgamma = st->invgamma; hst.p0_draw = st->p0_draw; hst.gamma_draw = st->gamma_draw; hst.nprop = st->nprop; }
    int rc = 0;

    for (int i = 0; i < N && rc == 0; ++i) {
        double *beta_i = betas + (size_t)i * p;
        const double tau2 = tau2s[i], sigma2 = sigma2s[i];
        for (int bi = 0; bi < nblock && rc == 0; ++bi) {
            const int32_t *blk = blocks + bi * bsize;
            for (int bj = 0; bj < bsize; ++bj) {                       /* zero the block :430-442 */
                const int j = blk[bj];
                const double old = beta_i[j];
                if (old != 0) {
                    const double *Xj = XT + (size_t)j * n;
                    for (int it = 0; it < n; ++it) r[it] += old * Xj[it];
                    if (probit == 1) for (int it = 0; it < n; ++it) mu_pred[it] += (-1 * old) * Xj[it];
                }
                beta_i[j] = 0;
            }
            for (int a = 0; a < bsize; ++a) {                          /* _precompute_suff_stats */
                for (int b = 0; b < bsize; ++b) XATXA_cache[a * bsize + b] = XTX[(size_t)blk[a] * p + blk[b]];
                const double *Xa = XT + (size_t)blk[a] * n;
                double s = 0.0;
                for (int it = 0; it < n; ++it) s += Xa[it] * r[it];
                XATr_cache[a] = s;
            }
            lprobs[0] = bsize * logp0;                                 /* :460 */
            for (int mj = 1; mj < nmodel; ++mj) {                      /* :461-489 */
                if (compute_QA(&w, masks[mj - 1], bsize, XATXA_cache, XATr_cache, tau2, sigma2)) { rc = 2; break; }
                double nrm = 0.0;
                for (int a = 0; a < w.m; ++a) nrm += w.mu[a] * w.mu[a];
                nrm = sqrt(nrm);                                       /* dnrm2 :479 */
                double e = tau2 * nrm * nrm / (2 * sigma2 * sigma2);   /* :480 */
                lprobs[mj] = e - w.logdet / 2;
                lprobs[mj] += (bsize - w.m) * logp0;
                lprobs[mj] += w.m * log1p0;
            }
            if (rc) break;
            /* _weighted_choice :55-79 */
            double maxlp = 0.0, denom = 0.0, cumsum = 0.0;
            for (int ii = 0; ii < nmodel; ++ii) maxlp = fmax(lprobs[ii], maxlp);
            for (int ii = 0; ii < nmodel; ++ii) { probs[ii] = exp(lprobs[ii] - maxlp); denom += probs[ii]; }
            const double u = (st && st->u) ? st->u[(size_t)i * nblock + bi]
                                           : orc_uniform(seed, chain, KIND_BLOCK, (uint32_t)bi, (uint32_t)i, 0);
            int mj = 0;
            for (int ii = 0; ii < nmodel; ++ii) {
                cumsum += probs[ii] / denom;
                if (cumsum >= u) { mj = ii; break; }
            }
            if (mj > 0) {                                              /* :535-884 */
                if (compute_QA(&w, masks[mj - 1], bsize, XATXA_cache, XATr_cache, tau2, sigma2)) { rc = 2; break; }
                const int m = w.m;
                const double cm = tau2 / sigma2;
                double mu[16], t1[16];
                for (int a = 0; a < m; ++a) {                          /* mu = Linv' (Linv XATr) :554-567 */
                    double s = 0.0;
                    for (int k = a; k < m; ++k) s += w.Linv[k * m + a] * w.mu[k];
                    mu[a] = s;
                }
                for (int a = 0; a < m; ++a) {                          /* cm * A * mu :571-586 */
                    double s = 0.0;
                    for (int k = 0; k < m; ++k) s += w.A[a * m + k] * mu[k];
                    t1[a] = cm * s;
                }
                for (int a = 0; a < m; ++a) { mu[a] = t1[a] - w.XATr[a]; mu[a] *= -1 * tau2 / sigma2; }   /* :587-597 */
                /* conditional covariance :599-754 (Fortran view): W1 = Linv A, W2 = Linv' W1, Vs = A W2 */
                double W1[256], W2[256], Vs[256], V[256];
                for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) {
                    double s = 0.0;
                    for (int k = 0; k <= a; ++k) s += w.Linv[a * m + k] * w.A[k * m + b];
                    W1[a * m + b] = s;
                }
                for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) {
                    double s = 0.0;
                    for (int k = a; k < m; ++k) s += w.Linv[k * m + a] * W1[k * m + b];
                    W2[a * m + b] = s;
                }
                for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) {
                    double s = 0.0;
                    for (int k = 0; k < m; ++k) s += w.A[a * m + k] * W2[k * m + b];
                    Vs[a * m + b] = s;
                }
                const double c2 = -1 * tau2 * tau2 / sigma2;
                const double c3 = tau2 * tau2 * tau2 / (sigma2 * sigma2);
                for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) {
                    double v = (a == b) ? tau2 : 0.0;
                    v += c2 * w.A[a * m + b];
                    v += c3 * Vs[a * m + b];
                    V[a * m + b] = v;
                }
                if (chol_lower(V, m, m)) { rc = 2; break; }            /* dpotrf :755-766 */
                double bn[16];
                for (int a = 0; a < m; ++a)
                    bn[a] = (st && st->z) ? st->z[((size_t)i * nblock + bi) * bsize + a]
                                          : orc_normal(seed, chain, KIND_BLOCK, (uint32_t)bi, (uint32_t)i, 1u + (uint32_t)a);
                for (int a = m - 1; a >= 0; --a) {                     /* beta = L_V z + mu :782-834 */
                    double s = 0.0;
                    for (int k = 0; k <= a; ++k) s += V[a * m + k] * bn[k];
                    bn[a] = s;
                }
                for (int a = 0; a < m; ++a) bn[a] += mu[a];
                for (int a = 0; a < m; ++a) {                          /* write back :850-873 */
                    const int j = blk[w.members[a]];
                    beta_i[j] = bn[a];
                    const double *Xj = XT + (size_t)j * n;
                    if (probit != 1) { const double neg = -1 * bn[a]; for (int it = 0; it < n; ++it) r[it] += neg * Xj[it]; }
                    else for (int it = 0; it < n; ++it) mu_pred[it] += beta_i[j] * Xj[it];
                }
                if (probit == 1) {                                     /* :875-884 */
                    for (int it = 0; it < n; ++it) {
                        double v = orc_truncnorm(seed, chain, event, (uint32_t)it, mu_pred[it], hp->sigma2, 0.0, (int)zlab[it]);
                        if (y_latent) y_latent[(size_t)i * n + it] = v;
                        y[it] = v;
                        r[it] = y[it] - mu_pred[it];
                    }
                    event++;
                }
            }
        }
        if (rc) break;
        orc_update_hparams(i, n, p, beta_i, r, hp, seed, chain, &hst, p0s, tau2s, sigma2s);
        logp0 = log(p0s[i]); log1p0 = log(1 - p0s[i]);                 /* :912-913 */
        if (i != N - 1) {                                              /* :914-918 */
            memcpy(betas + (size_t)(i + 1) * p, beta_i, sizeof(double) * p);
            p0s[i + 1] = p0s[i]; sigma2s[i + 1] = sigma2s[i]; tau2s[i + 1] = tau2s[i];
        }
    }
    if (n_events_out) *n_events_out = (int64_t)event;
    free(masks); free(XT); free(XTX); free(mu_pred); free(r); free(y); free(blocks); free(order);
    free(lprobs); free(probs);
    return rc;
}
