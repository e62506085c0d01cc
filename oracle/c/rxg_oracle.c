/* fp64 C twin of oracle/lgssm.py::smooth_reference_schedule -- TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: (1) a second, independently written restatement of the reference message schedule
 * (cross-checked against the NumPy oracle in tests/), and (2) the CPU baseline timed beside the
 * GPU kernels by bench.py (cpu_baseline.kind = "port"): it performs, per (chain, step), the same
 * 6 rule evaluations + 2 outbound products + 1 marginal the reference performs, in the same
 * parametrisations, with one SPD (Cholesky) inverse per mean-cov <-> weighted-mean-precision
 * conversion.  The reference itself (Julia +
 * ReactiveMP/ExponentialFamily/FastCholesky, /root/reference/Project.toml:43-73) cannot be
 * compiled here; this port is allocation-free and statically dispatched, i.e. an optimistic
 * stand-in for the reference's CPU speed.
 *
 * Layouts (batch innermost, as the C ABI):  y[T][m][batch] (fp32, upcast),
 * mean[T][d][batch], cov[T][d][d][batch] (fp64).  Model shared across chains.
 * Reference call sites: benchmarks/...Benchmark.ipynb:95-105 (model), src/inference/batch.jl:391-430
 * (iteration loop), src/model/plugins/reactivemp_inference.jl:365-374 (left-to-right product fold).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DMAX 64

typedef struct { int d; } dim_t;

/* C = A * B  (n x k)(k x p) */
static void mm(int n, int k, int p, const double* A, const double* B, double* C) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * p + j];
            C[i * p + j] = s;
        }
}
/* C = A' * B  (A is k x n) */
static void mtm(int n, int k, int p, const double* A, const double* B, double* C) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[l * n + i] * B[l * p + j];
            C[i * p + j] = s;
        }
}
/* C = A * B'  (B is p x k) */
static void mmt(int n, int k, int p, const double* A, const double* B, double* C) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
            C[i * p + j] = s;
        }
}
static void mv(int n, int k, const double* A, const double* x, double* y) {
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < k; ++l) s += A[i * k + l] * x[l];
        y[i] = s;
    }
}
static void mtv(int n, int k, const double* A, const double* x, double* y) { /* y = A' x, A is k x n */
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int l = 0; l < k; ++l) s += A[l * n + i] * x[l];
        y[i] = s;
    }
}

/* cholinv: SPD inverse via Cholesky (FastCholesky.cholinv restated).  Returns log det too.
 * 0 on success, 1 if not SPD. */
static int cholinv(int n, const double* S, double* Sinv, double* logdet, double* L, double* Li) {
    memset(L, 0, sizeof(double) * n * n);
    double ld = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = S[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        if (!(s > 0.0)) return 1;
        double ljj = sqrt(s);
        L[j * n + j] = ljj;
        ld += log(ljj);
        for (int i = j + 1; i < n; ++i) {
            double t = S[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / ljj;
        }
    }
    /* Li = inv(L) (lower) */
    memset(Li, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        Li[j * n + j] = 1.0 / L[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double t = 0.0;
            for (int k = j; k < i; ++k) t -= L[i * n + k] * Li[k * n + j];
            Li[i * n + j] = t / L[i * n + i];
        }
    }
    /* Sinv = Li' Li */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double t = 0.0;
            for (int k = i; k < n; ++k) t += Li[k * n + i] * Li[k * n + j];
            Sinv[i * n + j] = t;
            Sinv[j * n + i] = t;
        }
    if (logdet) *logdet = 2.0 * ld;
    return 0;
}

typedef struct {
    double *fwd_mu, *fwd_S;      /* [T][d], [T][d*d] forward (mu,Sigma) messages into x_t      */
    double *obs_xi;              /* [T][d]                                                      */
    double *t1, *t2, *t3, *L, *Li, *W, *Sg, *v1, *v2, *v3;
} work_t;

static work_t work_alloc(int T, int d, int m) {
    int n = d > m ? d : m;
    work_t w;
    w.fwd_mu = (double*)malloc(sizeof(double) * T * d);
    w.fwd_S = (double*)malloc(sizeof(double) * T * d * d);
    w.obs_xi = (double*)malloc(sizeof(double) * T * d);
    w.t1 = (double*)malloc(sizeof(double) * n * n); w.t2 = (double*)malloc(sizeof(double) * n * n);
    w.t3 = (double*)malloc(sizeof(double) * n * n); w.L = (double*)malloc(sizeof(double) * n * n);
    w.Li = (double*)malloc(sizeof(double) * n * n); w.W = (double*)malloc(sizeof(double) * n * n);
    w.Sg = (double*)malloc(sizeof(double) * n * n);
    w.v1 = (double*)malloc(sizeof(double) * n); w.v2 = (double*)malloc(sizeof(double) * n);
    w.v3 = (double*)malloc(sizeof(double) * n);
    return w;
}
static void work_free(work_t* w) {
    free(w->fwd_mu); free(w->fwd_S); free(w->obs_xi); free(w->t1); free(w->t2); free(w->t3);
    free(w->L); free(w->Li); free(w->W); free(w->Sg); free(w->v1); free(w->v2); free(w->v3);
}

/* One chain, reference schedule.  Returns number of non-SPD events (0 expected). */
static int smooth_chain(int d, int m, int T, long batch, long b,
                        const double* A, const double* B, const double* P, const double* Q,
                        const double* m0, const double* S0, const float* y,
                        double* mean, double* cov, double* nle_out, work_t* w) {
    int bad = 0;
    const int dd = d * d;
    /* observation message precision W^y = B' Q^-1 B is recomputed per step like the reference
     * (rule #3: (y,Q) -> weightedmean_precision: cholinv(Q); rule #4: B' . B). */
    double* Wy = (double*)malloc(sizeof(double) * dd);
    double* Qi = (double*)malloc(sizeof(double) * m * m);
    double* yv = (double*)malloc(sizeof(double) * m);
    double* xo = (double*)malloc(sizeof(double) * m);
    double* f_mu = (double*)malloc(sizeof(double) * d);
    double* f_S = (double*)malloc(sizeof(double) * dd);
    double nle = 0.0;

    memcpy(f_mu, m0, sizeof(double) * d);
    memcpy(f_S, S0, sizeof(double) * dd);                     /* rule #2, PointMass mean (prior)   */
    for (int t = 0; t < T; ++t) {
        memcpy(w->fwd_mu + (long)t * d, f_mu, sizeof(double) * d);
        memcpy(w->fwd_S + (long)t * dd, f_S, sizeof(double) * dd);
        for (int k = 0; k < m; ++k) yv[k] = (double)y[((long)t * m + k) * batch + b];
        /* rule #3 from data: (y, Q); to (xi,W): cholinv(Q) */
        bad += cholinv(m, Q, Qi, NULL, w->L, w->Li);
        mv(m, m, Qi, yv, xo);
        /* rule #4: (B' xi, B' W B) */
        mtv(d, m, B, xo, w->obs_xi + (long)t * d);
        mm(m, m, d, Qi, B, w->t1);
        mtm(d, m, d, B, w->t1, Wy);
        /* log evidence increment: innovation form */
        {
            double ld;
            mv(m, d, B, f_mu, w->v1);
            mm(m, d, d, B, f_S, w->t1);
            mmt(m, d, m, w->t1, B, w->t2);
            for (int i = 0; i < m * m; ++i) w->t2[i] += Q[i];
            bad += cholinv(m, w->t2, w->t3, &ld, w->L, w->Li);
            double q = 0.0;
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
                for (int j = 0; j < m; ++j) s += w->t3[i * m + j] * (yv[j] - w->v1[j]);
                q += (yv[i] - w->v1[i]) * s;
            }
            nle += 0.5 * (m * log(2.0 * M_PI) + ld + q);
        }
        /* product fwd x obs in (xi, W): cholinv(f_S) */
        bad += cholinv(d, f_S, w->W, NULL, w->L, w->Li);
        mv(d, d, w->W, f_mu, w->v1);
        for (int i = 0; i < d; ++i) w->v1[i] += w->obs_xi[(long)t * d + i];
        for (int i = 0; i < dd; ++i) w->W[i] += Wy[i];
        /* mean_cov of the filtered message for rule #1: cholinv(W) */
        bad += cholinv(d, w->W, w->Sg, NULL, w->L, w->Li);
        mv(d, d, w->Sg, w->v1, w->v2);
        /* rule #1: (A mu, A S A'), rule #2: + P */
        mv(d, d, A, w->v2, f_mu);
        mm(d, d, d, A, w->Sg, w->t1);
        mmt(d, d, d, w->t1, A, f_S);
        for (int i = 0; i < dd; ++i) f_S[i] += P[i];
    }

    /* backward */
    double* b_xi = (double*)calloc(d, sizeof(double));
    double* b_W = (double*)calloc(dd, sizeof(double));
    int have_bwd = 0;
    for (int t = T - 1; t >= 0; --t) {
        /* marginal: prod(fwd, obs, bwd) then mean_cov */
        bad += cholinv(d, w->fwd_S + (long)t * dd, w->W, NULL, w->L, w->Li);
        mv(d, d, w->W, w->fwd_mu + (long)t * d, w->v1);
        for (int i = 0; i < d; ++i) w->v1[i] += w->obs_xi[(long)t * d + i] + b_xi[i];
        for (int i = 0; i < dd; ++i) w->W[i] += Wy[i] + b_W[i];
        bad += cholinv(d, w->W, w->Sg, NULL, w->L, w->Li);
        mv(d, d, w->Sg, w->v1, w->v2);
        for (int i = 0; i < d; ++i) mean[((long)t * d + i) * batch + b] = w->v2[i];
        for (int i = 0; i < dd; ++i) cov[((long)t * dd + i) * batch + b] = w->Sg[i];
        if (t == 0) break;
        /* outbound of x_t toward its prior factor: prod(obs, bwd) -> (mu, Sigma): cholinv */
        for (int i = 0; i < d; ++i) w->v1[i] = w->obs_xi[(long)t * d + i] + b_xi[i];
        for (int i = 0; i < dd; ++i) w->W[i] = Wy[i] + b_W[i];
        bad += cholinv(d, w->W, w->Sg, NULL, w->L, w->Li);
        mv(d, d, w->Sg, w->v1, w->v2);
        /* rule #3': + P ; rule #4 needs (xi, W): cholinv */
        for (int i = 0; i < dd; ++i) w->Sg[i] += P[i];
        bad += cholinv(d, w->Sg, w->W, NULL, w->L, w->Li);
        mv(d, d, w->W, w->v2, w->v3);
        mtv(d, d, A, w->v3, b_xi);
        mm(d, d, d, w->W, A, w->t1);
        mtm(d, d, d, A, w->t1, b_W);
        have_bwd = 1;
    }
    (void)have_bwd;
    if (nle_out) nle_out[b] = nle;
    free(b_xi); free(b_W); free(Wy); free(Qi); free(yv); free(xo); free(f_mu); free(f_S);
    return bad;
}

/* Chains are processed in blocks of CB: the block's observations are gathered from the
 * batch-innermost ABI layout into a chain-major scratch copy, every chain runs on that contiguous
 * copy (so the per-thread working set stays in L2), and the block's posteriors are scattered back
 * with full-cache-line rows.  Without this the strided 8-byte accesses of the ABI layout make the
 * port memory-bound and it stops scaling beyond a few cores. */
#define CB 8
int rxo_lgssm_smooth_f64(int d, int m, int T, long batch,
                         const double* A, const double* B, const double* P, const double* Q,
                         const double* m0, const double* S0, const float* y,
                         double* mean, double* cov, double* nle, int nthreads) {
    if (d > DMAX || m > DMAX || d < 1 || m < 1 || T < 1 || batch < 1) return -1;
    int bad_total = 0;
    const long nblk = (batch + CB - 1) / CB;
    const int dd = d * d;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel reduction(+ : bad_total)
#endif
    {
        work_t w = work_alloc(T, d, m);
        float* yl = (float*)malloc(sizeof(float) * (size_t)T * m);          /* one chain, [T][m][1] */
        double* ml = (double*)malloc(sizeof(double) * (size_t)CB * T * d);   /* [c][T][d]            */
        double* cl = (double*)malloc(sizeof(double) * (size_t)CB * T * dd);  /* [c][T][d*d]          */
        double nl[CB];
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (long blk = 0; blk < nblk; ++blk) {
            const long b0 = blk * CB;
            const int nb = (int)((b0 + CB <= batch) ? CB : (batch - b0));
            for (int c = 0; c < nb; ++c) {
                for (long r = 0; r < (long)T * m; ++r) yl[r] = y[r * batch + b0 + c];
                bad_total += smooth_chain(d, m, T, 1, 0, A, B, P, Q, m0, S0, yl,
                                          ml + (size_t)c * T * d, cl + (size_t)c * T * dd, nl + c, &w);
            }
            for (long r = 0; r < (long)T * d; ++r)
                for (int c = 0; c < nb; ++c) mean[r * batch + b0 + c] = ml[(size_t)c * T * d + r];
            for (long r = 0; r < (long)T * dd; ++r)
                for (int c = 0; c < nb; ++c) cov[r * batch + b0 + c] = cl[(size_t)c * T * dd + r];
            if (nle) for (int c = 0; c < nb; ++c) nle[b0 + c] = nl[c];
        }
        free(yl); free(ml); free(cl);
        work_free(&w);
    }
    return bad_total;
}

int rxo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
