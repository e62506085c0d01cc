// Latent autoregressive model, structured VMP  q(x, x0) q(theta) q(gamma)  fused into one kernel: one thread = one series.
//
//     gamma ~ Gamma(a0, b0);  theta ~ N(0, I / w0);  x0 ~ N(0, I / p0)
//     x[t] ~ AR(x[t-1], theta, gamma)  (ARMeta(Multivariate | Univariate, order, ARsafe()));  y[t] ~ Normal(c'x[t], 1 / tau), c = e1
// [ref: /root/reference/test/models/autoregressive/lar_tests.jl:52-122 (model, constraints, initialisation, inference call);
//  the AR node's rules live in ReactiveMP (not vendored): DESIGN.md 3.11 restates them; the test suite pins the restatement
//  to the reference's free-energy values lar_tests.jl:170 (518.9182342, to 7e-8) and :201 (514.66086, inside its 0.01)].
//
// The AR factor is  N(y1 | theta'x, 1/gamma) * prod_{i>1} delta(y_i - x_{i-1}).  Under q(theta) q(gamma) the chain of x's is
// a linear-Gaussian state-space model with a companion transition A = companion(E theta), process noise 1/E[gamma] on the
// first coordinate only, a scalar observation of the first coordinate, and one extra factor exp(-1/2 E[gamma] x'V_theta x)
// on every x[t-1] (the AR(:y) / AR(:x) rules' "D = W_x + E[gamma] V_theta").  The reference runs it in information form
// with the deltas regularised by a precision of 1e12; here every VMP iteration is a covariance-form filter + RTS smoother in
// the exact limit (nothing is ever inverted that the limit makes singular):
//   forward, per step:  absorb the V_theta factor  (P' = P - P L (I + L'P L)^-1 L'P,  L L' = E[gamma] V_theta: one order x order
//                       Cholesky), predict through the companion matrix (O(order^2): a dot product and a shift), scalar update;
//                       (m', P') goes to the workspace [T][order + order(order+1)/2][batch] (coalesced over the batch);
//   backward, per step: RTS gain J = P'A'(A P'A' + V)^-1, smoothed (m, P) of x[t-1], Cov(x[t-1], x[t][1]); the sufficient
//                       statistics of the AR(:theta) and AR(:gamma) rules (sum of V_x + m_x m_x', V_y1x + m_x m_y1,
//                       V_y1 + m_y1^2) and the entropy / energy sums of the Bethe free energy in fp64 accumulators;
//   q(theta), q(gamma), free energy: O(order^3) in fp64.
// Bethe free energy (reference: AR average energy with the entropy correction for the degenerate coordinates, deterministic
// dot / * nodes, src/model/plugins/reactivemp_free_energy.jl:84-126), telescoped over the chain:
//   F = KL(q(theta)||p) + KL(q(gamma)||p) + E[-log p(x0)] - H[q(x0)] + sum_t (U_AR,t - H[x_t[1] | x_{t-1}]) + sum_t U_obs,t.
// Schedule per iteration (pinned by :170): q(x chain) from the previous q(theta), q(gamma); then q(theta); then q(gamma)
// with the new q(theta); free energy with everything new.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef RXG_HD
#define RXG_HD __host__ __device__ __forceinline__
#endif

namespace rxg {
namespace lar {

struct Params {
    float tau;               // observation precision
    float a0, b0;            // Gamma prior (shape, rate) on gamma
    float w0;                // theta prior precision (times identity), zero mean
    float p0;                // x0 prior precision (times identity), zero mean
    float init_shape, init_rate;   // initial q(gamma)
    float init_theta_prec;   // initial q(theta) = N(0, I / init_theta_prec)
};

// lower Cholesky factor of an SPD matrix; rd = reciprocal diagonal; a failed pivot is clamped and reported
template <typename S, int N>
RXG_HD void chol_factor(const S (&A)[N][N], S (&L)[N][N], S (&rd)[N], bool& bad) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        S s = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
        if (!(s > S(0))) { bad = true; s = S(1e-30); }
        const S d = sqrt(s);
        L[j][j] = d;
        rd[j] = S(1) / d;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            S t = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
            L[i][j] = t * rd[j];
        }
#pragma unroll
        for (int i = 0; i < j; ++i) L[i][j] = S(0);
    }
}
// Ai = A^-1 through the Cholesky factor; returns log det A
template <typename S, int N>
RXG_HD S spd_inverse(const S (&A)[N][N], S (&Ai)[N][N], bool& bad) {
    S L[N][N], rd[N], Li[N][N];
    chol_factor<S, N>(A, L, rd, bad);
    S logdet = S(0);
#pragma unroll
    for (int j = 0; j < N; ++j) logdet += log(L[j][j]);
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) Li[i][j] = S(0);
        Li[j][j] = rd[j];
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            S s = S(0);
#pragma unroll
            for (int k = j; k < i; ++k) s -= L[i][k] * Li[k][j];
            Li[i][j] = s * rd[i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            S s = S(0);
#pragma unroll
            for (int k = i; k < N; ++k) s += Li[k][i] * Li[k][j];
            Ai[i][j] = s;
            Ai[j][i] = s;
        }
    return S(2) * logdet;
}
// psi(x) in fp64: recurrence to x >= 6, then the asymptotic series
RXG_HD double digamma64(double x) {
    double dig = 0.0;
    while (x < 6.0) { dig -= 1.0 / x; x += 1.0; }
    const double i1 = 1.0 / x, i2 = i1 * i1;
    return dig + log(x) - 0.5 * i1 - i2 * (1.0 / 12.0 - i2 * (1.0 / 120.0 - i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0))));
}
// prediction through the companion matrix: mm = A m1, Pm = A P1 A' + diag(1/E[gamma], 0, ...)
template <int P>
RXG_HD void predict(const float (&th)[P], float inv_Eg, const float (&m1)[P], const float (&P1)[P][P], float (&mm)[P],
                    float (&Pm)[P][P], float (&Pth)[P]) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < P; ++k) s += P1[i][k] * th[k];
        Pth[i] = s;                                   // P1 theta = first column of P1 A'
        dot += th[i] * m1[i];
    }
    float q = inv_Eg;
#pragma unroll
    for (int i = 0; i < P; ++i) q += th[i] * Pth[i];
    mm[0] = dot;
    Pm[0][0] = q;
#pragma unroll
    for (int i = 1; i < P; ++i) {
        mm[i] = m1[i - 1];
        Pm[0][i] = Pth[i - 1];
        Pm[i][0] = Pth[i - 1];
#pragma unroll
        for (int j = 1; j < P; ++j) Pm[i][j] = P1[i - 1][j - 1];
    }
}

// One series, all iterations.  Layouts (batch innermost): y[T][batch]; ws[T][P + P(P+1)/2][batch];
// x_mean[T][P][batch], x_cov[T][P][P][batch] (last iteration; may be null); theta_mean[iters][P][batch],
// theta_cov[iters][P][P][batch], gamma_shape / gamma_rate[iters][batch], free_energy[iters][batch] (fp64; may be null).
template <int P>
RXG_HD bool chain(int64_t b, int64_t batch, const float* __restrict__ y, int T, int iters, const Params prm,
                  float* __restrict__ ws, float* __restrict__ x_mean, float* __restrict__ x_cov,
                  float* __restrict__ th_mean, float* __restrict__ th_cov, float* __restrict__ g_shape,
                  float* __restrict__ g_rate, double* __restrict__ fe) {
    constexpr int NS = P + P * (P + 1) / 2;
    constexpr double LOG2PI = 1.8378770664093453;
    bool bad = false;
    double mth[P], Vth[P][P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        mth[i] = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j) Vth[i][j] = (i == j) ? 1.0 / (double)prm.init_theta_prec : 0.0;
    }
    double ga = (double)prm.init_shape, gb = (double)prm.init_rate;
    const float inv_tau = 1.f / prm.tau;
    for (int it = 0; it < iters; ++it) {
        const double Egd = ga / gb;
        const float inv_Eg = (float)(1.0 / Egd);
        float th[P], Lg[P][P];
        {   // L L' = E[gamma] V_theta (fp64 factor, rounded once)
            double G[P][P], Ld[P][P], rdd[P];
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int j = 0; j < P; ++j) G[i][j] = Egd * Vth[i][j];
            chol_factor<double, P>(G, Ld, rdd, bad);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                th[i] = (float)mth[i];
#pragma unroll
                for (int j = 0; j < P; ++j) Lg[i][j] = (float)Ld[i][j];
            }
        }
        // ------------------------------------------------------------------ forward: filter
        float m[P], Pc[P][P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            m[i] = 0.f;
#pragma unroll
            for (int j = 0; j < P; ++j) Pc[i][j] = (i == j) ? 1.f / prm.p0 : 0.f;
        }
        for (int t = 0; t < T; ++t) {
            // absorb exp(-1/2 x' (L L') x):  S = I + L'P L = Ls Ls',  Z = P L Ls^-T,  P' = P - Z Z',  m' = m - Z Ls^-1 L'm
            float PL[P][P], S[P][P], Ls[P][P], rds[P], Z[P][P], w[P], m1[P], P1[P][P];
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int k = j; k < P; ++k) s += Pc[i][k] * Lg[k][j];
                    PL[i][j] = s;
                }
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    float s = (i == j) ? 1.f : 0.f;
#pragma unroll
                    for (int k = i; k < P; ++k) s += Lg[k][i] * PL[k][j];
                    S[i][j] = s;
                    S[j][i] = s;
                }
            chol_factor<float, P>(S, Ls, rds, bad);
#pragma unroll
            for (int r = 0; r < P; ++r)
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    float s = PL[r][j];
#pragma unroll
                    for (int k = 0; k < j; ++k) s -= Ls[j][k] * Z[r][k];
                    Z[r][j] = s * rds[j];
                }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                float s = 0.f;
#pragma unroll
                for (int k = i; k < P; ++k) s += Lg[k][i] * m[k];
                w[i] = s;
            }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                float s = w[i];
#pragma unroll
                for (int k = 0; k < i; ++k) s -= Ls[i][k] * w[k];
                w[i] = s * rds[i];
            }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                float s = m[i];
#pragma unroll
                for (int k = 0; k < P; ++k) s -= Z[i][k] * w[k];
                m1[i] = s;
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    float v = Pc[i][j];
#pragma unroll
                    for (int k = 0; k < P; ++k) v -= Z[i][k] * Z[j][k];
                    P1[i][j] = v;
                    P1[j][i] = v;
                }
            }
            {   // (m', P') -> workspace
                float* o = ws + ((int64_t)t * NS) * batch + b;
                int k = 0;
#pragma unroll
                for (int i = 0; i < P; ++i) o[(int64_t)(k++) * batch] = m1[i];
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) o[(int64_t)(k++) * batch] = P1[i][j];
            }
            float mm[P], Pm[P][P], Pth[P];
            predict<P>(th, inv_Eg, m1, P1, mm, Pm, Pth);
            const float yt = y[(int64_t)t * batch + b];
            const float is = 1.f / (Pm[0][0] + inv_tau);
            const float r = yt - mm[0];
            float kg[P];
#pragma unroll
            for (int i = 0; i < P; ++i) kg[i] = Pm[i][0] * is;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                m[i] = mm[i] + kg[i] * r;
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const float v = Pm[i][j] - kg[i] * Pm[0][j];
                    Pc[i][j] = v;
                    Pc[j][i] = v;
                }
            }
        }
        // ------------------------------------------------------------------ backward: RTS + statistics
        float ms[P], Ps[P][P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            ms[i] = m[i];
#pragma unroll
            for (int j = 0; j < P; ++j) Ps[i][j] = Pc[i][j];
        }
        double sC[P][P], sL[P], sR = 0.0, sLogS = 0.0, sObs = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            sL[i] = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) sC[i][j] = 0.0;
        }
        const bool last = (it == iters - 1);
        for (int t = T - 1; t >= 0; --t) {
            if (last && x_mean) {
#pragma unroll
                for (int i = 0; i < P; ++i) x_mean[((int64_t)t * P + i) * batch + b] = ms[i];
            }
            if (last && x_cov) {
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j < P; ++j) x_cov[(((int64_t)t * P + i) * P + j) * batch + b] = Ps[i][j];
            }
            const float yt = y[(int64_t)t * batch + b];
            const float r = yt - ms[0];
            sObs += (double)(r * r + Ps[0][0]);
            float m1[P], P1[P][P];
            {
                const float* o = ws + ((int64_t)t * NS) * batch + b;
                int k = 0;
#pragma unroll
                for (int i = 0; i < P; ++i) m1[i] = o[(int64_t)(k++) * batch];
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) {
                        const float v = o[(int64_t)(k++) * batch];
                        P1[i][j] = v;
                        P1[j][i] = v;
                    }
            }
            float mm[P], Pm[P][P], Pth[P], Pmi[P][P], J[P][P];
            predict<P>(th, inv_Eg, m1, P1, mm, Pm, Pth);
            spd_inverse<float, P>(Pm, Pmi, bad);
            // J = (P1 A') Pm^-1,  (P1 A')[i][0] = (P1 theta)[i],  (P1 A')[i][j] = P1[i][j-1]
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    float s = Pth[i] * Pmi[0][j];
#pragma unroll
                    for (int k = 1; k < P; ++k) s += P1[i][k - 1] * Pmi[k][j];
                    J[i][j] = s;
                }
            float mprev[P], Pprev[P][P], cross[P], JD[P][P];
#pragma unroll
            for (int i = 0; i < P; ++i) {
                float s = m1[i], cs = 0.f;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    s += J[i][k] * (ms[k] - mm[k]);
                    cs += J[i][k] * Ps[k][0];
                }
                mprev[i] = s;
                cross[i] = cs;                           // Cov(x[t-1][i], x[t][1])
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < P; ++k) v += J[i][k] * (Ps[k][j] - Pm[k][j]);
                    JD[i][j] = v;
                }
            }
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    float v = P1[i][j];
#pragma unroll
                    for (int k = 0; k < P; ++k) v += JD[i][k] * J[j][k];
                    Pprev[i][j] = v;
                    Pprev[j][i] = v;
                }
            {   // conditional variance of x[t][1] given x[t-1] under the smoothed joint
                float Lp[P][P], rdp[P], wv[P];
                chol_factor<float, P>(Pprev, Lp, rdp, bad);
                float sc = Ps[0][0];
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    float s = cross[i];
#pragma unroll
                    for (int k = 0; k < i; ++k) s -= Lp[i][k] * wv[k];
                    wv[i] = s * rdp[i];
                    sc -= wv[i] * wv[i];
                }
                if (!(sc > 0.f)) { bad = true; sc = 1e-30f; }
                sLogS += (double)logf(sc);
            }
            sR += (double)(Ps[0][0] + ms[0] * ms[0]);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                sL[i] += (double)(cross[i] + mprev[i] * ms[0]);
#pragma unroll
                for (int j = 0; j <= i; ++j) sC[i][j] += (double)(Pprev[i][j] + mprev[i] * mprev[j]);
            }
#pragma unroll
            for (int i = 0; i < P; ++i) {
                ms[i] = mprev[i];
#pragma unroll
                for (int j = 0; j < P; ++j) Ps[i][j] = Pprev[i][j];
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = i + 1; j < P; ++j) sC[i][j] = sC[j][i];
        // ------------------------------------------------------------------ q(theta), q(gamma), free energy (fp64)
        double W[P][P], logdetW;
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j < P; ++j) W[i][j] = Egd * sC[i][j] + ((i == j) ? (double)prm.w0 : 0.0);
        logdetW = spd_inverse<double, P>(W, Vth, bad);
        double Bsum = sR, trV = 0.0, mmth = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) s += Vth[i][j] * (Egd * sL[j]);
            mth[i] = s;
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            Bsum -= 2.0 * mth[i] * sL[i];
            trV += Vth[i][i];
            mmth += mth[i] * mth[i];
#pragma unroll
            for (int j = 0; j < P; ++j) Bsum += mth[i] * sC[i][j] * mth[j] + Vth[i][j] * sC[j][i];
        }
        ga = (double)prm.a0 + 0.5 * T;
        gb = (double)prm.b0 + 0.5 * Bsum;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            th_mean[((int64_t)it * P + i) * batch + b] = (float)mth[i];
#pragma unroll
            for (int j = 0; j < P; ++j) th_cov[(((int64_t)it * P + i) * P + j) * batch + b] = (float)Vth[i][j];
        }
        g_shape[(int64_t)it * batch + b] = (float)ga;
        g_rate[(int64_t)it * batch + b] = (float)gb;
        if (fe) {
            const double dig = digamma64(ga), Elog = dig - log(gb), Egn = ga / gb;
            const double a0 = (double)prm.a0, b0 = (double)prm.b0, w0 = (double)prm.w0, p0 = (double)prm.p0;
            const double U_ar = 0.5 * T * (LOG2PI - Elog) + 0.5 * Egn * Bsum;
            const double H_cond = 0.5 * T * (1.0 + LOG2PI) + 0.5 * sLogS;
            const double U_obs = 0.5 * T * (LOG2PI - log((double)prm.tau)) + 0.5 * (double)prm.tau * sObs;
            double V0[P][P], V0i[P][P], tr0 = 0.0, mm0 = 0.0;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                tr0 += (double)Ps[i][i];
                mm0 += (double)ms[i] * (double)ms[i];
#pragma unroll
                for (int j = 0; j < P; ++j) V0[i][j] = (double)Ps[i][j];
            }
            const double logdetV0 = spd_inverse<double, P>(V0, V0i, bad);
            const double U_x0 = 0.5 * (P * LOG2PI - P * log(p0) + p0 * (tr0 + mm0));
            const double H_x0 = 0.5 * (P * (1.0 + LOG2PI) + logdetV0);
            const double kl_t = 0.5 * (w0 * (trV + mmth) - P - P * log(w0) + logdetW);     // log det V_theta = -log det W
            const double kl_g = (ga - a0) * dig - lgamma(ga) + lgamma(a0) + a0 * (log(gb) - log(b0)) + ga * (b0 - gb) / gb;
            fe[(int64_t)it * batch + b] = kl_t + kl_g + U_x0 - H_x0 + U_ar - H_cond + U_obs;
        }
    }
    return bad;
}

}  // namespace lar
}  // namespace rxg
