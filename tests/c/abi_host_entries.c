/* Plain-C host of librxgauss, second program: the entries bound in round 2, driven the way a ccall / cgo binding without a
 * CUDA runtime of its own would drive them -- device buffers from rxg_device_alloc, data moved with rxg_memcpy_h2d / _d2h,
 * options through rxg_set_option.  TEST CODE.  No oracle here: every check is a self-consistency property of the call
 * (round trips, filter == smoother at the last step, free energies that must not increase, documented output identities);
 * parity against the fp64 oracle of the same entries is tests/test_*_gpu.py.
 *
 *   gcc -std=c99 -O2 -Iinclude tests/c/abi_host_entries.c -o tests/c/abi_host_entries \
 *       -Lrxinfer.jl_b200 -lrxgauss -Wl,-rpath,'$ORIGIN/../../rxinfer.jl_b200' -lm
 * Exit code 0 = all properties hold; 2 = no CUDA device (there is no CPU fallback); 1 = a property failed.          */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rxgauss.h"

static unsigned long long lcg = 0x9E3779B97F4A7C15ULL;
static double urand(void) {
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((lcg >> 11) + 0.5) / 9007199254740992.0;
}
static double nrand(void) { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }

static rxg_ctx* ctx;
static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#define CALL(x) do { int rc_ = (x); if (rc_ != RXG_OK) { fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, rxg_last_error(ctx)); return 1; } } while (0)

static float* dev(size_t n) {
    void* p = NULL;
    if (rxg_device_alloc(ctx, n * sizeof(float), &p) != RXG_OK) { fprintf(stderr, "rxg_device_alloc: %s\n", rxg_last_error(ctx)); exit(1); }
    return (float*)p;
}
static float* up(const float* h, size_t n) {
    float* d = dev(n);
    if (rxg_memcpy_h2d(ctx, d, h, n * sizeof(float)) != RXG_OK) { fprintf(stderr, "rxg_memcpy_h2d: %s\n", rxg_last_error(ctx)); exit(1); }
    return d;
}
static void down(void* h, const void* d, size_t bytes) {
    if (rxg_memcpy_d2h(ctx, h, d, bytes) != RXG_OK) { fprintf(stderr, "rxg_memcpy_d2h: %s\n", rxg_last_error(ctx)); exit(1); }
}
static double rel(const float* a, const float* b, size_t n) {
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) { num += ((double)a[i] - b[i]) * ((double)a[i] - b[i]); den += (double)b[i] * b[i]; }
    return sqrt(num / (den > 0 ? den : 1));
}

/* mean_cov -> weightedmean_precision -> mean_cov is the identity; prod(q, q) doubles (xi, W)  (d = 4: register-resident
 * kernels, d = 16: shared-memory kernels) */
static int rules_round_trip(int d, int n) {
    const size_t nv = (size_t)d * n, nm = (size_t)d * d * n;
    float *mu = malloc(nv * 4), *S = malloc(nm * 4), *mu2 = malloc(nv * 4), *S2 = malloc(nm * 4), *W2 = malloc(nm * 4), *W = malloc(nm * 4);
    double* X = malloc(sizeof(double) * d * d);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < d * d; ++k) X[k] = nrand();
        for (int r = 0; r < d; ++r) {
            mu[(size_t)r * n + i] = (float)nrand();
            for (int c = 0; c < d; ++c) {
                double s = (r == c) ? d : 0.0;
                for (int k = 0; k < d; ++k) s += X[r * d + k] * X[c * d + k];
                S[((size_t)r * d + c) * n + i] = (float)s;
            }
        }
    }
    float *dmu = up(mu, nv), *dS = up(S, nm), *dxi = dev(nv), *dW = dev(nm), *dmu2 = dev(nv), *dS2 = dev(nm), *dxi2 = dev(nv), *dW2 = dev(nm);
    int32_t* dst = (int32_t*)dev(n);
    int32_t* st = malloc(sizeof(int32_t) * n);
    CALL(rxg_meancov_to_wmp_f32(ctx, n, d, dmu, dS, dxi, dW, dst, RXG_PTR_DEVICE));
    CALL(rxg_wmp_to_meancov_f32(ctx, n, d, dxi, dW, dmu2, dS2, dst, RXG_PTR_DEVICE));
    CALL(rxg_prod_gaussian_f32(ctx, n, d, dxi, dW, dxi, dW, dxi2, dW2, RXG_PTR_DEVICE));
    down(mu2, dmu2, nv * 4); down(S2, dS2, nm * 4); down(W, dW, nm * 4); down(W2, dW2, nm * 4); down(st, dst, sizeof(int32_t) * n);
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += st[i] != RXG_OK;
    for (size_t i = 0; i < nm; ++i) W[i] *= 2.f;
    const double em = rel(mu2, mu, nv), es = rel(S2, S, nm), ep = rel(W2, W, nm);
    printf("rules d=%d: round trip mean %.2e cov %.2e, prod(q, q) = 2 W: %.2e, non-SPD flags %d\n", d, em, es, ep, bad);
    CHECK(em < 1e-4 && es < 1e-4 && ep < 1e-6 && bad == 0, "rule round trip d=%d", d);
    rxg_device_free(ctx, dmu); rxg_device_free(ctx, dS); rxg_device_free(ctx, dxi); rxg_device_free(ctx, dW); rxg_device_free(ctx, dmu2);
    rxg_device_free(ctx, dS2); rxg_device_free(ctx, dxi2); rxg_device_free(ctx, dW2); rxg_device_free(ctx, dst);
    free(mu); free(S); free(mu2); free(S2); free(W2); free(W); free(X); free(st);
    return 0;
}

/* filtered and smoothed posteriors coincide at the last step, and both calls report the same evidence; a (d, m) = (5, 3)
 * model has no dedicated kernel family (rxg_supports still says yes: embedded) */
static int filter_vs_smoother(void) {
    enum { D = 5, M = 3, T = 40, BATCH = 64 };
    float A[D * D] = {0}, B[M * D], P[D * D] = {0}, Q[M * M] = {0}, S0[D * D] = {0}, m0[D] = {0};
    for (int i = 0; i < D; ++i) { A[i * D + i] = 0.9f; if (i + 1 < D) A[i * D + i + 1] = 0.2f; P[i * D + i] = 0.3f; S0[i * D + i] = 4.f; }
    for (int i = 0; i < M * D; ++i) B[i] = (float)(0.5 * nrand());
    for (int i = 0; i < M; ++i) Q[i * M + i] = 1.5f;
    CHECK(rxg_supports(D, M) == 1 && rxg_supports(65, 1) == 0, "rxg_supports");
    float* y = malloc(sizeof(float) * T * M * BATCH);
    for (size_t i = 0; i < (size_t)T * M * BATCH; ++i) y[i] = (float)(2.0 * nrand());
    float *dy = up(y, (size_t)T * M * BATCH), *dms = dev((size_t)T * D * BATCH), *dcs = dev((size_t)T * D * D * BATCH), *dns = dev(BATCH);
    float *dmf = dev((size_t)T * D * BATCH), *dcf = dev((size_t)T * D * D * BATCH), *dnf = dev(BATCH);
    CALL(rxg_lgssm_smooth_f32(ctx, D, M, T, BATCH, A, B, P, Q, m0, S0, NULL, dy, NULL, dms, dcs, dns, NULL, RXG_PTR_DEVICE));
    CALL(rxg_lgssm_filter_f32(ctx, D, M, T, BATCH, A, B, P, Q, m0, S0, NULL, dy, NULL, dmf, dcf, dnf, NULL, RXG_PTR_DEVICE));
    float ms[D * BATCH], mf[D * BATCH], cs[D * D * BATCH], cf[D * D * BATCH], ns[BATCH], nf[BATCH];
    down(ms, dms + (size_t)(T - 1) * D * BATCH, sizeof ms); down(mf, dmf + (size_t)(T - 1) * D * BATCH, sizeof mf);
    down(cs, dcs + (size_t)(T - 1) * D * D * BATCH, sizeof cs); down(cf, dcf + (size_t)(T - 1) * D * D * BATCH, sizeof cf);
    down(ns, dns, sizeof ns); down(nf, dnf, sizeof nf);
    const double em = rel(ms, mf, D * BATCH), ec = rel(cs, cf, D * D * BATCH), en = rel(ns, nf, BATCH);
    printf("lgssm (5,3): smoother vs filter at the last step: mean %.2e cov %.2e evidence %.2e\n", em, ec, en);
    CHECK(em < 1e-5 && ec < 1e-5 && en < 1e-6, "filter vs smoother");
    rxg_device_free(ctx, dy); rxg_device_free(ctx, dms); rxg_device_free(ctx, dcs); rxg_device_free(ctx, dns);
    rxg_device_free(ctx, dmf); rxg_device_free(ctx, dcf); rxg_device_free(ctx, dnf);
    free(y);
    return 0;
}

/* the parameter-learning models: free energies must not increase (beyond round-off), documented output identities hold */
static int vmp_models(void) {
    enum { T = 300, BATCH = 48, ITERS = 8 };
    float* s = malloc(sizeof(float) * T * BATCH);
    for (int b = 0; b < BATCH; ++b) {           /* AR(2) series, observed with noise */
        double x1 = 0, x2 = 0;
        for (int t = 0; t < T; ++t) {
            const double x = 0.6 * x1 - 0.3 * x2 + 0.5 * nrand();
            x2 = x1; x1 = x;
            s[(size_t)t * BATCH + b] = (float)(x + 0.3 * nrand());
        }
    }
    float* ds = up(s, (size_t)T * BATCH);
    /* --- autoregressive regression (ar_tests.jl) */
    {
        const int p = 2;
        float *tm = dev(p * BATCH), *tc = dev(p * p * BATCH), *gs = dev(BATCH), *gr = dev(BATCH);
        double* fe = (double*)dev(2 * ITERS * BATCH);
        CALL(rxg_ar_vmp_f32(ctx, p, T, BATCH, ITERS, 1.f, 1.f, 1.f, 1.f, 1.f, ds, tm, tc, gs, gr, fe, RXG_PTR_DEVICE));
        double h[ITERS * BATCH]; float th[2 * BATCH];
        down(h, fe, sizeof h); down(th, tm, sizeof th);
        int up_ = 0;
        for (int it = 1; it < ITERS; ++it) for (int b = 0; b < BATCH; ++b) up_ += h[it * BATCH + b] > h[(it - 1) * BATCH + b] + 1e-6;
        printf("ar_vmp: free energy %.4f -> %.4f, theta[0] = (%.3f, %.3f)\n", h[0], h[(ITERS - 1) * BATCH], th[0], th[BATCH]);
        CHECK(up_ == 0 && fabs(th[0] - 0.6f) < 0.4f, "ar_vmp");   /* attenuated by the observation noise (errors in variables): 0.39 */
        rxg_device_free(ctx, tm); rxg_device_free(ctx, tc); rxg_device_free(ctx, gs); rxg_device_free(ctx, gr); rxg_device_free(ctx, fe);
    }
    /* --- latent autoregressive model (lar_tests.jl) */
    {
        const int p = 2;
        const float prm[8] = {10.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        float *xm = dev((size_t)T * p * BATCH), *xc = dev((size_t)T * p * p * BATCH), *tm = dev(ITERS * p * BATCH), *tc = dev(ITERS * p * p * BATCH);
        float *gs = dev(ITERS * BATCH), *gr = dev(ITERS * BATCH);
        double* fe = (double*)dev(2 * ITERS * BATCH);
        int32_t* st = (int32_t*)dev(BATCH);
        CALL(rxg_lar_vmp_f32(ctx, p, T, BATCH, ITERS, prm, ds, xm, xc, tm, tc, gs, gr, fe, st, RXG_PTR_DEVICE));
        double h[ITERS * BATCH]; int32_t hs[BATCH]; float x0[BATCH];
        down(h, fe, sizeof h); down(hs, st, sizeof hs); down(x0, xm + (size_t)(T - 1) * p * BATCH, sizeof x0);
        int up_ = 0, bad = 0; double dev_ = 0;
        for (int b = 0; b < BATCH; ++b) { bad += hs[b] != 0; up_ += !(h[(ITERS - 1) * BATCH + b] < h[b]); dev_ += fabs(x0[b] - s[(size_t)(T - 1) * BATCH + b]); }
        printf("lar_vmp: free energy %.4f -> %.4f, mean |x_T[1] - y_T| = %.3f\n", h[0], h[(ITERS - 1) * BATCH], dev_ / BATCH);
        CHECK(up_ == 0 && bad == 0 && dev_ / BATCH < 0.5, "lar_vmp");
        CHECK(rxg_lar_vmp_f32(ctx, 7, T, BATCH, ITERS, prm, ds, xm, xc, tm, tc, gs, gr, fe, st, RXG_PTR_DEVICE) == RXG_ERR_UNSUPPORTED, "lar_vmp order 7 must be refused");
        rxg_device_free(ctx, xm); rxg_device_free(ctx, xc); rxg_device_free(ctx, tm); rxg_device_free(ctx, tc);
        rxg_device_free(ctx, gs); rxg_device_free(ctx, gr); rxg_device_free(ctx, fe); rxg_device_free(ctx, st);
    }
    /* --- Gamma-precision VMP around the scalar smoother, with its free energy */
    {
        float *pm = dev((size_t)T * BATCH), *pv = dev((size_t)T * BATCH), *sh = dev(BATCH), *ra = dev(BATCH), *fe = dev(ITERS * BATCH);
        CALL(rxg_lgssm_vmp_gamma_fe_f32(ctx, T, BATCH, ITERS, 1.f, 1.f, 0.f, 100.f, 1.f, 1.f, 1.f, ds, pm, pv, sh, ra, fe, RXG_PTR_DEVICE));
        float h[ITERS * BATCH], hsh[BATCH];
        down(h, fe, sizeof h); down(hsh, sh, sizeof hsh);
        int up_ = 0;
        for (int it = 1; it < ITERS; ++it) for (int b = 0; b < BATCH; ++b) up_ += h[it * BATCH + b] > h[(it - 1) * BATCH + b] + 1e-3f + 2e-6f * fabsf(h[(it - 1) * BATCH + b]);
        printf("lgssm_vmp_gamma_fe: free energy %.3f -> %.3f, shape = %.1f (a0 + T/2 = %.1f)\n", h[0], h[(ITERS - 1) * BATCH], hsh[0], 1.0 + T / 2.0);
        CHECK(up_ == 0 && fabs(hsh[0] - (1.0 + T / 2.0)) < 1e-3, "lgssm_vmp_gamma_fe");
        rxg_device_free(ctx, pm); rxg_device_free(ctx, pv); rxg_device_free(ctx, sh); rxg_device_free(ctx, ra); rxg_device_free(ctx, fe);
    }
    /* --- HGF with its free energy: variances positive, free energy finite */
    {
        const float init[4] = {0.f, 5.f, 0.f, 5.f};
        float *out = dev((size_t)T * 4 * BATCH), *fe = dev((size_t)T * 4 * BATCH);
        CALL(rxg_hgf_filter_fe_f32(ctx, T, BATCH, 4, 1.f, 0.f, 0.04f, 0.01f, init, NULL, ds, out, fe, RXG_PTR_DEVICE));
        float* h = malloc(sizeof(float) * T * 4 * BATCH); float* hf = malloc(sizeof(float) * T * 4 * BATCH);
        down(h, out, sizeof(float) * T * 4 * BATCH); down(hf, fe, sizeof(float) * T * 4 * BATCH);
        int bad = 0;
        for (int t = 0; t < T; ++t) for (int b = 0; b < BATCH; ++b) {
            bad += !(h[((size_t)t * 4 + 1) * BATCH + b] > 0.f) || !(h[((size_t)t * 4 + 3) * BATCH + b] > 0.f);
            for (int it = 0; it < 4; ++it) bad += !isfinite(hf[((size_t)t * 4 + it) * BATCH + b]);
        }
        printf("hgf_filter_fe: %d bad entries, last free energy %.4f\n", bad, hf[((size_t)(T - 1) * 4 + 3) * BATCH]);
        CHECK(bad == 0, "hgf_filter_fe");
        rxg_device_free(ctx, out); rxg_device_free(ctx, fe); free(h); free(hf);
    }
    rxg_device_free(ctx, ds);
    free(s);
    return 0;
}

int main(void) {
    int rc = rxg_create(&ctx, 0, 0);
    if (rc == RXG_ERR_NO_DEVICE) { fprintf(stderr, "no CUDA device: librxgauss has no CPU fallback\n"); return 2; }
    if (rc != RXG_OK) { fprintf(stderr, "rxg_create -> %d\n", rc); return 1; }
    long long v = -1;
    CHECK(rxg_set_option(ctx, RXG_OPT_HOST_THREADS, 3) == RXG_OK && rxg_get_option(ctx, RXG_OPT_HOST_THREADS, &v) == RXG_OK && v == 3, "option round trip");
    CHECK(rxg_set_option(ctx, RXG_OPT_COUNT_, 1) != RXG_OK, "unknown option must be refused");
    rxg_set_option(ctx, RXG_OPT_HOST_THREADS, 0);
    if (rules_round_trip(4, 1001) || rules_round_trip(16, 203) || filter_vs_smoother() || vmp_models()) return 1;
    rxg_destroy(ctx);
    if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    printf("abi_host_entries: all properties hold\n");
    return 0;
}
