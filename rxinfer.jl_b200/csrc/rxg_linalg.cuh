// Small dense linear algebra on register-resident matrices (compile-time shapes, fully unrolled).
// Used by every kernel family of librxgauss; S = float for the per-chain sweeps, double for the
// gain-table precompute.  These are the device twins of the d x d operations the reference does
// through Julia LinearAlgebra / FastCholesky.cholinv (re-exported at
// /root/reference/src/RxInfer.jl:6) inside every @rule body.
#pragma once
#include <cuda_runtime.h>

namespace rxg {

template <typename S, int R, int C>
struct Mat {
    S a[R * C];
    __device__ __forceinline__ S& operator()(int i, int j) { return a[i * C + j]; }
    __device__ __forceinline__ const S& operator()(int i, int j) const { return a[i * C + j]; }
};
template <typename S, int N>
struct Vec {
    S a[N];
    __device__ __forceinline__ S& operator()(int i) { return a[i]; }
    __device__ __forceinline__ const S& operator()(int i) const { return a[i]; }
};

template <typename S> __device__ __forceinline__ S fma_(S a, S b, S c);
template <> __device__ __forceinline__ float fma_<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fma_<double>(double a, double b, double c) { return fma(a, b, c); }
template <typename S> __device__ __forceinline__ S rsqrt_(S x);
template <> __device__ __forceinline__ float rsqrt_<float>(float x) {
    // rsqrtf is ~2 ulp; one Newton step brings the Cholesky pivots to fp32 round-off
    float r = rsqrtf(x);
    return r * __fmaf_rn(-0.5f * x * r, r, 1.5f);
}
template <> __device__ __forceinline__ double rsqrt_<double>(double x) { return 1.0 / sqrt(x); }
template <typename S> __device__ __forceinline__ S log_(S x);
template <> __device__ __forceinline__ float log_<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double log_<double>(double x) { return log(x); }

// C = A * B            (R x K)(K x C)
template <typename S, int R, int K, int C>
__device__ __forceinline__ Mat<S, R, C> mul(const Mat<S, R, K>& A, const Mat<S, K, C>& B) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            S s = A(i, 0) * B(0, j);
#pragma unroll
            for (int k = 1; k < K; ++k) s = fma_<S>(A(i, k), B(k, j), s);
            o(i, j) = s;
        }
    return o;
}
// C = A * B'           (R x K)(C x K)'
template <typename S, int R, int K, int C>
__device__ __forceinline__ Mat<S, R, C> mul_nt(const Mat<S, R, K>& A, const Mat<S, C, K>& B) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            S s = A(i, 0) * B(j, 0);
#pragma unroll
            for (int k = 1; k < K; ++k) s = fma_<S>(A(i, k), B(j, k), s);
            o(i, j) = s;
        }
    return o;
}
// C = A' * B           (K x R)'(K x C)
template <typename S, int R, int K, int C>
__device__ __forceinline__ Mat<S, R, C> mul_tn(const Mat<S, K, R>& A, const Mat<S, K, C>& B) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            S s = A(0, i) * B(0, j);
#pragma unroll
            for (int k = 1; k < K; ++k) s = fma_<S>(A(k, i), B(k, j), s);
            o(i, j) = s;
        }
    return o;
}
// symmetric product  T * A' + Add  where the result is known symmetric (T = A * Ssym):
// computes the lower triangle and mirrors it.  T is N x K, A is N x K.
template <typename S, int N, int K>
__device__ __forceinline__ Mat<S, N, N> sym_mul_nt_add(const Mat<S, N, K>& T, const Mat<S, N, K>& A,
                                                       const Mat<S, N, N>& Add) {
    Mat<S, N, N> o;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            S s = Add(i, j);
#pragma unroll
            for (int k = 0; k < K; ++k) s = fma_<S>(T(i, k), A(j, k), s);
            o(i, j) = s;
            o(j, i) = s;
        }
    return o;
}
// Base - V V'  (symmetric downdate), V is N x K
template <typename S, int N, int K>
__device__ __forceinline__ Mat<S, N, N> sym_downdate(const Mat<S, N, N>& Base, const Mat<S, N, K>& V) {
    Mat<S, N, N> o;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            S s = Base(i, j);
#pragma unroll
            for (int k = 0; k < K; ++k) s = fma_<S>(-V(i, k), V(j, k), s);
            o(i, j) = s;
            o(j, i) = s;
        }
    return o;
}
template <typename S, int R, int C>
__device__ __forceinline__ Vec<S, R> mulv(const Mat<S, R, C>& A, const Vec<S, C>& x) {
    Vec<S, R> o;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        S s = A(i, 0) * x(0);
#pragma unroll
        for (int k = 1; k < C; ++k) s = fma_<S>(A(i, k), x(k), s);
        o(i) = s;
    }
    return o;
}
// y = A' x    (A is K x R)
template <typename S, int R, int K>
__device__ __forceinline__ Vec<S, R> mulv_t(const Mat<S, K, R>& A, const Vec<S, K>& x) {
    Vec<S, R> o;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        S s = A(0, i) * x(0);
#pragma unroll
        for (int k = 1; k < K; ++k) s = fma_<S>(A(k, i), x(k), s);
        o(i) = s;
    }
    return o;
}

// Cholesky factor with reciprocal diagonal.  L(i,i) holds 1 / l_ii (NOT l_ii); strict lower part
// holds l_ij; upper part is zero.  `bad` is set when a pivot is not strictly positive (or NaN) --
// the pivot is then clamped so the sweep can continue and the chain is reported through status[].
// logdet accumulates log det S = -2 sum log(1 / l_ii) when WANT_LOGDET.
template <typename S, int N>
struct Chol {
    Mat<S, N, N> L;
    S neg_half_logdet;  // sum_i log(1/l_ii) = -0.5 log det
};
template <typename S, int N, bool WANT_LOGDET>
__device__ __forceinline__ Chol<S, N> cholesky(const Mat<S, N, N>& A, bool& bad) {
    Chol<S, N> c;
    c.neg_half_logdet = S(0);
#pragma unroll
    for (int i = 0; i < N * N; ++i) c.L.a[i] = S(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        S s = A(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma_<S>(-c.L(j, k), c.L(j, k), s);
        if (!(s > S(0))) { bad = true; s = S(1e-30); }
        const S r = rsqrt_<S>(s);
        c.L(j, j) = r;
        if (WANT_LOGDET) c.neg_half_logdet += log_<S>(r);
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            S t = A(i, j);
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma_<S>(-c.L(i, k), c.L(j, k), t);
            c.L(i, j) = t * r;
        }
    }
    return c;
}
// z = L^-1 e  (forward substitution; L in the reciprocal-diagonal form above)
template <typename S, int N>
__device__ __forceinline__ Vec<S, N> solve_L(const Mat<S, N, N>& L, const Vec<S, N>& e) {
    Vec<S, N> z;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        S s = e(i);
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma_<S>(-L(i, k), z(k), s);
        z(i) = s * L(i, i);
    }
    return z;
}
// x = L^-T z (back substitution)
template <typename S, int N>
__device__ __forceinline__ Vec<S, N> solve_Lt(const Mat<S, N, N>& L, const Vec<S, N>& z) {
    Vec<S, N> x;
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        S s = z(i);
#pragma unroll
        for (int k = i + 1; k < N; ++k) s = fma_<S>(-L(k, i), x(k), s);
        x(i) = s * L(i, i);
    }
    return x;
}
// V = Y L^-T   (each row r: solve L v = y_r), Y is R x N
template <typename S, int R, int N>
__device__ __forceinline__ Mat<S, R, N> solve_right_Lt(const Mat<S, R, N>& Y, const Mat<S, N, N>& L) {
    Mat<S, R, N> V;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            S s = Y(r, j);
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma_<S>(-L(j, k), V(r, k), s);
            V(r, j) = s * L(j, j);
        }
    return V;
}
// G = U L^-1   (each row r: solve g L = u_r), U is R x N
template <typename S, int R, int N>
__device__ __forceinline__ Mat<S, R, N> solve_right_L(const Mat<S, R, N>& U, const Mat<S, N, N>& L) {
    Mat<S, R, N> G;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            S s = U(r, j);
#pragma unroll
            for (int k = j + 1; k < N; ++k) s = fma_<S>(-G(r, k), L(k, j), s);
            G(r, j) = s * L(j, j);
        }
    return G;
}
// SPD inverse through Cholesky: inv(A) = L^-T L^-1  (cholinv)
template <typename S, int N>
__device__ __forceinline__ Mat<S, N, N> cholinv(const Mat<S, N, N>& A, bool& bad) {
    Chol<S, N> c = cholesky<S, N, false>(A, bad);
    // Li = L^-1 (lower), column by column
    Mat<S, N, N> Li;
#pragma unroll
    for (int i = 0; i < N * N; ++i) Li.a[i] = S(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        Li(j, j) = c.L(j, j);
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            S s = S(0);
#pragma unroll
            for (int k = j; k < i; ++k) s = fma_<S>(-c.L(i, k), Li(k, j), s);
            Li(i, j) = s * c.L(i, i);
        }
    }
    Mat<S, N, N> o;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            S s = S(0);
#pragma unroll
            for (int k = i; k < N; ++k) s = fma_<S>(Li(k, i), Li(k, j), s);
            o(i, j) = s;
            o(j, i) = s;
        }
    return o;
}

template <typename S, int R, int C>
__device__ __forceinline__ Mat<S, R, C> add(const Mat<S, R, C>& A, const Mat<S, R, C>& B) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] + B.a[i];
    return o;
}
template <typename S, int N>
__device__ __forceinline__ Mat<S, N, N> identity() {
    Mat<S, N, N> o;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) o(i, j) = (i == j) ? S(1) : S(0);
    return o;
}
template <typename S, typename S2, int R, int C>
__device__ __forceinline__ Mat<S, R, C> convert(const Mat<S2, R, C>& A) {
    Mat<S, R, C> o;
#pragma unroll
    for (int i = 0; i < R * C; ++i) o.a[i] = (S)A.a[i];
    return o;
}

}  // namespace rxg
