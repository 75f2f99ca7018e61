/*
 * ocbnn_oracle.c — plain C restatement (fp32, closed-form gradients; single-threaded, callers fan chains out over host threads) of the OC-BNN log posterior,
 * its gradient and the HMC iteration.  TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/_ref/libocbnn_oracle.so
 * and used by tests/ (cross-check against the numpy oracle and the golden vectors of the unmodified reference) and by
 * bench.py's CPU legs (cpu_baseline / --impl reference) as the multi-core CPU baseline.  Never linked or loaded by the product.
 *
 * Follows the reference (dtak/ocbnn-public):
 *   forward                     bnn/base.py:72-104        Gaussian prior (isotropic / AOCP)   base.py:106-133
 *   Gaussian likelihood         base.py:252-265           softmax likelihood                  base.py:432-444
 *   Bernoulli likelihood        base.py:524-542 (naive sigmoid e^z/(e^z+1))
 *   negative exponential COCP   base.py:360-388 (tanh form; the joint gap decision is made by the caller)
 *   positive Gaussian COCP      base.py:344-358 (K = 1)  positive Dirichlet COCP             base.py:498-511
 *   HMC leapfrog + MH           bnn/inference.py:29-78 (no restore on reject)
 * Pinned by tests/test_oracle_c.py against tests/golden/ (generated from the reference, see oracle/make_golden.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OC_MAXL 8
#define OC_MAXW 256

typedef struct {
    int n_layers;
    int sizes[OC_MAXL + 1];
    int act;   /* 0 identity, 1 rbf, 2 relu */
    int task;  /* 0 regression, 1 classifier, 2 binary */
    float sigma_w, sigma_noise;
    const float *prior_mean, *prior_std; /* per weight, or NULL for N(0, sigma_w) */
    int n_train;
    const float *x, *y;
    int n_neg, maxg;
    const float *neg_x;
    const int *neg_ngaps;
    const float *neg_gaps; /* n_neg x maxg x 2 */
    float neg_scale, tau0, tau1;
    int n_pos;
    const float *pos_x, *pos_y;
    float sigma_c;
    int n_dir;
    const float *dir_x, *dir_cm1, *dir_norm; /* cm1: n_dir x sK ; norm: lgamma terms per point */
} oc_spec;

static int oc_nweights(const oc_spec *s) {
    int d = 0;
    for (int l = 1; l <= s->n_layers; ++l) d += (s->sizes[l - 1] + 1) * s->sizes[l];
    return d;
}

static float act_f(int act, float z) { return act == 1 ? expf(-z * z) : act == 2 ? (z > 0 ? z : 0) : z; }
static float dact_f(int act, float z, float a) { return act == 1 ? a * (-2.f * z) : act == 2 ? (z > 0 ? 1.f : 0.f) : 1.f; }

/* Phi(z) = 1/4 (tanh(-tau0 z)+1)(tanh(-tau1 z)+1) and its derivative (base.py:360-362) */
static void phi(const oc_spec *s, float z, float *P, float *D) {
    float t0 = tanhf(-s->tau0 * z), t1 = tanhf(-s->tau1 * z);
    *P = 0.25f * (t0 + 1.f) * (t1 + 1.f);
    *D = 0.25f * (-s->tau0 * (1.f - t0 * t0) * (t1 + 1.f) - s->tau1 * (1.f - t1 * t1) * (t0 + 1.f));
}

/* forward + backward of one point; df (sK) is produced by the head given f; returns the head value */
typedef float (*head_fn)(const oc_spec *, const float *f, const void *ctx, int idx, float *df);

static float point_term(const oc_spec *s, const float *w, const float *x, head_fn head, const void *ctx, int idx, float scale, float *g) {
    float a[OC_MAXL + 1][OC_MAXW], z[OC_MAXL + 1][OC_MAXW], dz[OC_MAXW], dn[OC_MAXW];
    const int nl = s->n_layers;
    for (int d = 0; d < s->sizes[0]; ++d) a[0][d] = x[d];
    int off = 0, offs[OC_MAXL + 1];
    for (int l = 1; l <= nl; ++l) {
        const int si = s->sizes[l - 1], so = s->sizes[l];
        offs[l] = off;
        for (int o = 0; o < so; ++o) {
            float acc = w[off + si * so + o];
            for (int d = 0; d < si; ++d) acc += a[l - 1][d] * w[off + d * so + o];
            z[l][o] = acc;
            a[l][o] = l < nl ? act_f(s->act, acc) : acc;
        }
        off += (si + 1) * so;
    }
    const float val = head(s, z[nl], ctx, idx, dz);
    if (g) {
        for (int o = 0; o < s->sizes[nl]; ++o) dz[o] *= scale;
        for (int l = nl; l >= 1; --l) {
            const int si = s->sizes[l - 1], so = s->sizes[l];
            const float *wl = w + offs[l];
            float *gl = g + offs[l];
            for (int d = 0; d < si; ++d)
                for (int o = 0; o < so; ++o) gl[d * so + o] += a[l - 1][d] * dz[o];
            for (int o = 0; o < so; ++o) gl[si * so + o] += dz[o];
            if (l > 1) {
                for (int d = 0; d < si; ++d) {
                    float acc = 0.f;
                    for (int o = 0; o < so; ++o) acc += wl[d * so + o] * dz[o];
                    dn[d] = acc * dact_f(s->act, z[l - 1][d], a[l - 1][d]);
                }
                memcpy(dz, dn, sizeof(float) * si);
            }
        }
    }
    return scale * val;
}

static float head_lik(const oc_spec *s, const float *f, const void *ctx, int n, float *df) {
    const int sk = s->sizes[s->n_layers];
    (void)ctx;
    if (s->task == 0) {
        const float sn = s->sigma_noise;
        float v = 0.f;
        if (sk == 1) {
            const float r = f[0] - s->y[n];
            df[0] = -r / (sn * sn);
            return -(r * r) / (2.f * sn * sn) - logf(sn) - 0.91893853f;
        }
        for (int k = 0; k < sk; ++k) {
            const float r = f[k] - s->y[n * sk + k];
            v += -0.5f * r * r / sn;
            df[k] = -r / sn;
        }
        return v - 0.5f * sk * (1.8378770664f + logf(sn));
    }
    if (s->task == 1) {
        float mx = f[0], se = 0.f;
        for (int k = 1; k < sk; ++k) mx = f[k] > mx ? f[k] : mx;
        for (int k = 0; k < sk; ++k) se += expf(f[k] - mx);
        const int yl = (int)s->y[n];
        for (int k = 0; k < sk; ++k) df[k] = (k == yl ? 1.f : 0.f) - expf(f[k] - mx) / se;
        return f[yl] - mx - logf(se);
    }
    {   /* naive sigmoid of the reference; t log p + (1-t) log(1-p) */
        const float t = s->y[n], e = expf(f[0]), p = e / (e + 1.0f);
        df[0] = t - p;
        return t * logf(p) + (1.0f - t) * logf(1.f - p);
    }
}

static float head_neg(const oc_spec *s, const float *f, const void *ctx, int c, float *df) {
    (void)ctx;
    float pen = 0.f, d = 0.f;
    for (int k = 0; k < s->neg_ngaps[c]; ++k) {
        const float a = s->neg_gaps[(c * s->maxg + k) * 2], b = s->neg_gaps[(c * s->maxg + k) * 2 + 1];
        float p1, d1, p2, d2;
        phi(s, a - f[0], &p1, &d1);
        phi(s, f[0] - b, &p2, &d2);
        pen += p1 * p2;
        d += -d1 * p2 + p1 * d2;
    }
    df[0] = d;
    return pen;
}

static float head_pos(const oc_spec *s, const float *f, const void *ctx, int c, float *df) {
    (void)ctx;
    const float r = f[0] - s->pos_y[c];
    df[0] = -r / s->sigma_c;
    return -0.5f * r * r / s->sigma_c - 0.5f * logf(6.2831853f * s->sigma_c);
}

static float head_dir(const oc_spec *s, const float *f, const void *ctx, int c, float *df) {
    (void)ctx;
    const int sk = s->sizes[s->n_layers];
    float mx = f[0], se = 0.f, tot = 0.f, v = 0.f;
    for (int k = 1; k < sk; ++k) mx = f[k] > mx ? f[k] : mx;
    for (int k = 0; k < sk; ++k) se += expf(f[k] - mx);
    for (int k = 0; k < sk; ++k) tot += s->dir_cm1[c * sk + k];
    for (int k = 0; k < sk; ++k) {
        const float p = expf(f[k] - mx) / se;
        v += s->dir_cm1[c * sk + k] * logf(p);
        df[k] = s->dir_cm1[c * sk + k] - tot * p;
    }
    return v + s->dir_norm[c];
}

static float logpost_one(const oc_spec *s, const float *w, int with_data, float *g) {
    const int d = oc_nweights(s);
    float lp = 0.f;
    if (g) memset(g, 0, sizeof(float) * d);
    for (int i = 0; i < d; ++i) {
        const float mu = s->prior_mean ? s->prior_mean[i] : 0.f, sd = s->prior_std ? s->prior_std[i] : s->sigma_w;
        const float df = w[i] - mu;
        lp += -(df * df) / (2.f * sd * sd) - logf(sd) - 0.91893853f;
        if (g) g[i] += -df / (sd * sd);
    }
    const int s0 = s->sizes[0];
    for (int c = 0; c < s->n_pos; ++c) lp += point_term(s, w, s->pos_x + c * s0, head_pos, 0, c, 1.f, g);
    for (int c = 0; c < s->n_neg; ++c) lp += point_term(s, w, s->neg_x + c * s0, head_neg, 0, c, s->neg_scale, g);
    for (int c = 0; c < s->n_dir; ++c) lp += point_term(s, w, s->dir_x + c * s0, head_dir, 0, c, 1.f, g);
    if (with_data)
        for (int n = 0; n < s->n_train; ++n) lp += point_term(s, w, s->x + n * s0, head_lik, 0, n, 1.f, g);
    return lp;
}

void oc_logpost_grad(const oc_spec *s, const float *w, int m, int with_data, float *lp, float *grad) {
    const int d = oc_nweights(s);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < m; ++c) lp[c] = logpost_one(s, w + (long)c * d, with_data, grad ? grad + (long)c * d : 0);
}

/* n_it HMC iterations of m chains (inference.py:39-78); the proposal is kept on reject like the reference */
void oc_hmc_iterate(const oc_spec *s, float *q, const float *p0, const double *u, int m, int n_it, float eps, int L, int with_data,
                    unsigned char *acc) {
    const int d = oc_nweights(s);
    const float eh = (float)((double)eps / 2.0);
#pragma omp parallel
    {
        float *g = (float *)malloc(sizeof(float) * d), *p = (float *)malloc(sizeof(float) * d);
#pragma omp for schedule(static)
        for (int c = 0; c < m; ++c) {
            float *qc = q + (long)c * d;
            for (int it = 0; it < n_it; ++it) {
                memcpy(p, p0 + ((long)it * m + c) * d, sizeof(float) * d);
                const float U0 = -logpost_one(s, qc, with_data, 0);
                float K0 = 0.f;
                for (int i = 0; i < d; ++i) K0 += p[i] * p[i];
                K0 *= 0.5f;
                logpost_one(s, qc, with_data, g);
                for (int i = 0; i < d; ++i) p[i] += eh * g[i];
                for (int l = 1; l <= L; ++l) {
                    for (int i = 0; i < d; ++i) qc[i] += eps * p[i];
                    if (l < L) {
                        logpost_one(s, qc, with_data, g);
                        for (int i = 0; i < d; ++i) p[i] += eps * g[i];
                    }
                }
                const float U1 = -logpost_one(s, qc, with_data, g);
                for (int i = 0; i < d; ++i) p[i] += eh * g[i];
                float K1 = 0.f;
                for (int i = 0; i < d; ++i) K1 += p[i] * p[i];
                K1 *= 0.5f;
                const double alpha = (double)expf(U0 - U1 + K0 - K1);
                if (acc) acc[(long)it * m + c] = u[(long)it * m + c] < alpha;
            }
        }
        free(g);
        free(p);
    }
}

/* ---- SVGD (inference.py:208-259), row ranges so that the caller can fan rows out over host threads -------------------- */
/* distances ||x_i - x_j + 1e-6|| of the pairs i < j with i in [row_begin, row_end), appended to out; returns their number */
long oc_svgd_pair_dists(const float *x, int n, int d, int row_begin, int row_end, float *out) {
    long k = 0;
    for (int i = row_begin; i < row_end; ++i)
        for (int j = i + 1; j < n; ++j) {
            float acc = 0.f;
            for (int c = 0; c < d; ++c) {
                const float e = x[(long)i * d + c] - x[(long)j * d + c] + 1e-6f;
                acc += e * e;
            }
            out[k++] = sqrtf(acc);
        }
    return k;
}

/* phi_i = (1/n) sum_j [K_ji g_j - (2/h) K_ji (x_j - x_i)],  K_ji = exp(-||x_j - x_i||^2 / h), rows i in [row_begin, row_end) */
void oc_svgd_phi_rows(const float *x, const float *g, int n, int d, float h, int row_begin, int row_end, float *phi) {
    const float c2 = 2.f / h, invn = 1.f / (float)n;
    float *acc = (float *)malloc(sizeof(float) * d);
    for (int i = row_begin; i < row_end; ++i) {
        memset(acc, 0, sizeof(float) * d);
        const float *xi = x + (long)i * d;
        for (int j = 0; j < n; ++j) {
            const float *xj = x + (long)j * d, *gj = g + (long)j * d;
            float d2 = 0.f;
            for (int c = 0; c < d; ++c) {
                const float e = xj[c] - xi[c];
                d2 += e * e;
            }
            const float k = expf(-d2 / h);
            for (int c = 0; c < d; ++c) acc[c] += k * (gj[c] - c2 * (xj[c] - xi[c]));
        }
        for (int c = 0; c < d; ++c) phi[(long)(i - row_begin) * d + c] = acc[c] * invn;
    }
    free(acc);
}
