/* Plain-C host of librxgauss: what a non-Python binding (Julia ccall, C, Fortran ...) sees.  TEST CODE.
 *
 * Generates a small batched LGSSM (notebook model lifted to d = 4, SURVEY.md 8d config 1/2 shape), runs
 *   rxg_lgssm_smooth_f32   with HOST pointers (staged path)   and
 *   rxg_lgssm_filter_chunk_f32 / device pointers are NOT used here (no CUDA headers in this file),
 * and checks posterior means / covariances / evidence against the fp64 C twin of the reference schedule
 * (oracle/c/rxg_oracle.c, linked in: test infrastructure checking the product, never the other way round).
 * Exit code 0 = parity within the contract tolerances (mean rel-L2 < 1e-5, cov rel-F < 1e-4, evidence rel < 1e-5).
 *
 *   gcc -std=c99 -O2 -Iinclude tests/c/abi_host.c oracle/c/rxg_oracle.c -o tests/c/abi_host \
 *       -Lrxinfer.jl_b200 -lrxgauss -Wl,-rpath,'$ORIGIN/../../rxinfer.jl_b200' -lm -fopenmp
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rxgauss.h"

int rxo_lgssm_smooth_f64(int d, int m, int T, long batch, const double* A, const double* B, const double* P,
                         const double* Q, const double* m0, const double* S0, const float* y, double* mean,
                         double* cov, double* nle, int nthreads);

static unsigned long long lcg = 88172645463325252ULL;
static double urand(void) {
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((lcg >> 11) + 0.5) / 9007199254740992.0;
}
static double nrand(void) { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }

int main(void) {
    enum { D = 4, T = 64, BATCH = 96 };
    const double PI = 3.14159265358979323846;
    double A64[D * D] = {0}, B64[D * D] = {0}, P64[D * D] = {0}, Q64[D * D] = {0}, S064[D * D] = {0}, m064[D] = {0};
    const double th[2] = {PI / 15, PI / 35}, bdiag[D] = {1.3, 0.7, 1.3, 0.7};
    for (int b = 0; b < 2; ++b) {
        A64[(2 * b) * D + 2 * b] = cos(th[b]);     A64[(2 * b) * D + 2 * b + 1] = -sin(th[b]);
        A64[(2 * b + 1) * D + 2 * b] = sin(th[b]); A64[(2 * b + 1) * D + 2 * b + 1] = cos(th[b]);
    }
    for (int i = 0; i < D; ++i) { B64[i * D + i] = bdiag[i]; P64[i * D + i] = 0.05; Q64[i * D + i] = 10.0; S064[i * D + i] = 100.0; }
    float A[D * D], B[D * D], P[D * D], Q[D * D], S0[D * D], m0[D];
    for (int i = 0; i < D * D; ++i) {   /* the ABI sees the model rounded to fp32 once; the oracle gets the same values */
        A[i] = (float)A64[i]; B[i] = (float)B64[i]; P[i] = (float)P64[i]; Q[i] = (float)Q64[i]; S0[i] = (float)S064[i];
        A64[i] = A[i]; B64[i] = B[i]; P64[i] = P[i]; Q64[i] = Q[i]; S064[i] = S0[i];
    }
    for (int i = 0; i < D; ++i) m0[i] = 0.f;

    float* y = NULL; float* mean = NULL; float* cov = NULL; float* nle = NULL;
    if (rxg_host_alloc((void**)&y, sizeof(float) * T * D * BATCH) != RXG_OK ||      /* pinned: asynchronous staging */
        rxg_host_alloc((void**)&mean, sizeof(float) * T * D * BATCH) != RXG_OK ||
        rxg_host_alloc((void**)&cov, sizeof(float) * T * D * D * BATCH) != RXG_OK ||
        rxg_host_alloc((void**)&nle, sizeof(float) * BATCH) != RXG_OK) {
        fprintf(stderr, "rxg_host_alloc failed (no CUDA device?)\n");
        return 2;
    }
    for (int b = 0; b < BATCH; ++b) {     /* x_t = A x_{t-1} + N(0,P), y_t = B x_t + N(0,Q)  (ipynb:134-148) */
        double x[D] = {0, 0, 0, 0};
        for (int t = 0; t < T; ++t) {
            double xn[D];
            for (int i = 0; i < D; ++i) {
                double s = 0;
                for (int j = 0; j < D; ++j) s += A64[i * D + j] * x[j];
                xn[i] = s + sqrt(0.05) * nrand();
            }
            for (int i = 0; i < D; ++i) { x[i] = xn[i]; y[((size_t)t * D + i) * BATCH + b] = (float)(bdiag[i] * x[i] + sqrt(10.0) * nrand()); }
        }
    }
    rxg_ctx* ctx = NULL;
    int rc = rxg_create(&ctx, 0, 0);
    if (rc != RXG_OK) { fprintf(stderr, "rxg_create -> %d (the hot path has no CPU fallback)\n", rc); return 2; }
    rc = rxg_lgssm_smooth_f32(ctx, D, D, T, BATCH, A, B, P, Q, m0, S0, NULL, y, NULL, mean, cov, nle, NULL, 0u);
    if (rc != RXG_OK) { fprintf(stderr, "rxg_lgssm_smooth_f32 -> %d: %s\n", rc, rxg_last_error(ctx)); return 1; }

    double* rm = malloc(sizeof(double) * T * D * BATCH);
    double* rc64 = malloc(sizeof(double) * T * D * D * BATCH);
    double* rn = malloc(sizeof(double) * BATCH);
    if (rxo_lgssm_smooth_f64(D, D, T, BATCH, A64, B64, P64, Q64, m064, S064, y, rm, rc64, rn, 1) != 0) return 1;
    double em = 0, nm = 0, ec = 0, nc = 0, en = 0;
    for (size_t i = 0; i < (size_t)T * D * BATCH; ++i) { em += (mean[i] - rm[i]) * (mean[i] - rm[i]); nm += rm[i] * rm[i]; }
    for (size_t i = 0; i < (size_t)T * D * D * BATCH; ++i) { ec += (cov[i] - rc64[i]) * (cov[i] - rc64[i]); nc += rc64[i] * rc64[i]; }
    for (int b = 0; b < BATCH; ++b) { const double e = fabs(nle[b] - rn[b]) / fabs(rn[b]); if (e > en) en = e; }
    em = sqrt(em / nm); ec = sqrt(ec / nc);
    printf("abi_host: launches=%lld mean relL2=%.3e cov relF=%.3e evidence rel=%.3e\n", rxg_launch_count(ctx), em, ec, en);
    rxg_destroy(ctx);
    rxg_host_free(y); rxg_host_free(mean); rxg_host_free(cov); rxg_host_free(nle);
    return (em < 1e-5 && ec < 1e-4 && en < 1e-5) ? 0 : 1;
}
