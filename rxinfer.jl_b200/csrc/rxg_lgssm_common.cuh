// Definitions shared by the LGSSM sweep kernels (rxg_lgssm.cu) and the gain-table precompute
// (rxg_gain.cuh): model parameter block, gain-table record layout, small load/store helpers.
#pragma once
#include "rxg_internal.h"
#include "rxg_linalg.cuh"

namespace rxg {

template <int D, int M>
struct ModelF {
    float A[D * D], B[M * D], P[D * D], Q[M * M], m0[D], S0[D * D];
    float u[D];   // constant transition offset: x[t] ~ N(A x[t-1] + u, P)  (the `+` rule with a PointMass operand)
};

struct PerChainPtrs {
    const float *A, *B, *P, *Q, *m0, *S0, *u;
};

template <typename S, int R, int C>
__device__ __forceinline__ Mat<S, R, C> load_const(const float* p) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R * C; ++i) o.a[i] = (S)p[i];
    return o;
}
template <typename S, int R, int C>
__device__ __forceinline__ Mat<S, R, C> load_strided(const float* p, int64_t stride) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R * C; ++i) o.a[i] = (S)__ldg(p + i * stride);
    return o;
}
template <typename S, int R, int C>
__device__ __forceinline__ Mat<S, C, R> transpose(const Mat<S, R, C>& A) {
    Mat<S, C, R> o;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) o(j, i) = A(i, j);
    return o;
}

#define RXG_HALF_LOG_2PI 0.91893853320467274178

constexpr int pad4(int n) { return (n + 3) / 4 * 4; }

template <int D, int M>
struct Tab {
    // forward record, per t
    static constexpr int F_OFF = 0;                        // (I - K B) A           D x D
    static constexpr int K_OFF = F_OFF + pad4(D * D);      // Kalman gain           D x M
    static constexpr int LI_OFF = K_OFF + pad4(D * M);     // L^-1 of innovation    M x M (lower)
    static constexpr int C_OFF = LI_OFF + pad4(M * M);     // M/2 log 2pi + 1/2 log det S
    static constexpr int GF_OFF = C_OFF + 4;               // (I - K B) u            D
    static constexpr int FWD_REC = GF_OFF + pad4(D);
    // backward record, per t
    static constexpr int E_OFF = 0;                        // I - G A               D x D
    static constexpr int G_OFF = E_OFF + pad4(D * D);      // RTS gain              D x D
    static constexpr int SS_OFF = G_OFF + pad4(D * D);     // smoothed covariance   D x D
    static constexpr int GB_OFF = SS_OFF + pad4(D * D);    // -G u                   D
    static constexpr int BWD_REC = GB_OFF + pad4(D);
    static constexpr int SF_REC = pad4(D * D);             // filtered covariance   D x D
};

struct GainWs {
    float* fwd;    // [T][FWD_REC]
    float* bwd;    // [T][BWD_REC]
    float* sf;     // [T][SF_REC]
    double* Sp;    // [T][D*D] predicted covariance
    double* Sf;    // [T][D*D] filtered covariance
    double* Cc;    // [T][D*D] conditional covariance Sf - U U'
    double* Gd;    // [T][D*D] RTS gain (fp64)
};

template <int R, int C>
__device__ __forceinline__ void store_d(double* p, const Mat<double, R, C>& A) {
#pragma unroll
    for (int i = 0; i < R * C; ++i) p[i] = A.a[i];
}
template <int R, int C>
__device__ __forceinline__ Mat<double, R, C> load_d(const double* p) {
    Mat<double, R, C> o;
#pragma unroll
    for (int i = 0; i < R * C; ++i) o.a[i] = p[i];
    return o;
}
template <int N>
__device__ __forceinline__ void store_fv(float* p, const Vec<double, N>& v) {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = (float)v.a[i];
}
template <int R, int C>
__device__ __forceinline__ void store_f(float* p, const Mat<double, R, C>& A) {
#pragma unroll
    for (int i = 0; i < R * C; ++i) p[i] = (float)A.a[i];
}

}  // namespace rxg
