"""Development prototype (NumPy) of the covariance-form schedule rxg_lar_vmp runs on the device: used to fix the
algebra and to see what float32 recursions cost in accuracy before writing CUDA.  Not shipped, not imported."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import vmp
from scipy.special import digamma, gammaln


def chol_inv(M):
    return np.linalg.inv(M.astype(np.float64)).astype(M.dtype)


def run(y, p, tau, iters, dt):
    T = len(y)
    y = y.astype(dt)
    mth = np.zeros(p, dt); Vth = np.eye(p, dtype=dt)
    ga, gb = 1.0, 1.0
    a0 = b0 = 1.0
    fes = []
    for _ in range(iters):
        Eg = dt(ga / gb)
        G = Eg * Vth
        L = np.linalg.cholesky(G.astype(np.float64)).astype(dt)
        A = np.zeros((p, p), dt); A[0] = mth
        for i in range(1, p): A[i, i - 1] = 1
        m = np.zeros(p, dt); P = np.eye(p, dtype=dt)
        mp_s = np.zeros((T + 1, p), dt); Pp_s = np.zeros((T + 1, p, p), dt)   # (m', P') stored per AR_t (index t, about x_{t-1})
        for t in range(1, T + 1):
            # G factor on x_{t-1}
            PL = P @ L
            S = np.eye(p, dtype=dt) + L.T @ PL
            Si = chol_inv(S)
            K = PL @ Si
            m1 = m - K @ (L.T @ m)
            P1 = P - K @ PL.T
            P1 = dt(0.5) * (P1 + P1.T)
            mp_s[t] = m1; Pp_s[t] = P1
            # predict
            mm = A @ m1
            Pm = A @ P1 @ A.T; Pm[0, 0] += dt(1) / Eg
            # scalar observation of coordinate 0
            s = Pm[0, 0] + dt(1 / tau)
            k = Pm[:, 0] / s
            m = mm + k * (y[t - 1] - mm[0])
            P = Pm - np.outer(k, Pm[0, :])
            P = dt(0.5) * (P + P.T)
        ms = m.copy(); Ps = P.copy()            # smoothed x_T
        sC = np.zeros((p, p)); sL = np.zeros(p); sR = 0.0; sH = 0.0; sU = 0.0
        xm = np.zeros((T, p), dt); xc = np.zeros((T, p, p), dt)
        for t in range(T, 0, -1):
            xm[t - 1] = ms; xc[t - 1] = Ps
            sU += 0.5 * (np.log(2 * np.pi) - np.log(tau) + tau * (float(y[t - 1] - ms[0]) ** 2 + float(Ps[0, 0])))
            m1 = mp_s[t]; P1 = Pp_s[t]
            mm = A @ m1
            Pm = A @ P1 @ A.T; Pm[0, 0] += dt(1) / Eg
            J = P1 @ A.T @ chol_inv(Pm)
            mprev = m1 + J @ (ms - mm)
            Pprev = P1 + J @ (Ps - Pm) @ J.T
            Pprev = dt(0.5) * (Pprev + Pprev.T)
            cross = (J @ Ps)[:, 0]                     # Cov(x_{t-1}, x_t[0])
            Vy1 = Ps[0, 0]; my1 = ms[0]
            z = chol_inv(Pprev) @ cross
            s_cond = Vy1 - cross @ z
            sH += 0.5 * (1 + np.log(2 * np.pi) + np.log(float(s_cond)))
            sC += (Pprev + np.outer(mprev, mprev)).astype(np.float64)
            sL += (cross + mprev * my1).astype(np.float64)
            sR += float(Vy1 + my1 * my1)
            ms, Ps = mprev, Pprev
        # x0 marginal = (ms, Ps)
        W = np.eye(p) + float(Eg) * sC
        Vth64 = np.linalg.inv(W); mth64 = Vth64 @ (float(Eg) * sL)
        Bsum = sR - 2 * mth64 @ sL + mth64 @ sC @ mth64 + np.trace(Vth64 @ sC)
        ga = a0 + 0.5 * T; gb = b0 + 0.5 * Bsum
        Egn = ga / gb; Elog = digamma(ga) - np.log(gb)
        U_ar = 0.5 * T * (np.log(2 * np.pi) - Elog) + 0.5 * Egn * Bsum
        m0 = ms.astype(np.float64); V0 = Ps.astype(np.float64)
        U_x0 = 0.5 * (p * np.log(2 * np.pi) + np.trace(V0) + m0 @ m0)
        H_x0 = 0.5 * (p * (1 + np.log(2 * np.pi)) + np.linalg.slogdet(V0)[1])
        kl_t = 0.5 * (np.trace(Vth64) + mth64 @ mth64 - p - np.linalg.slogdet(Vth64)[1])
        kl_g = (ga - a0) * digamma(ga) - gammaln(ga) + gammaln(a0) + a0 * (np.log(gb) - np.log(b0)) + ga * (b0 - gb) / gb
        fes.append(kl_t + kl_g + U_x0 - H_x0 + U_ar - sH + sU)
        mth = mth64.astype(dt); Vth = Vth64.astype(dt)
    return np.array(fes), xm, xc, mth64, Vth64, ga, gb


if __name__ == "__main__":
    st, obs = vmp.latent_ar_reference_data()
    for p in (1, 2, 5):
        ref = vmp.latent_ar(obs[:, None], p, 5.0, 15)
        for dt in (np.float64, np.float32):
            fe, xm, xc, mth, Vth, ga, gb = run(obs, p, 5.0, 15, dt)
            rfe = ref["free_energy"][:, 0]
            print(p, dt.__name__, "fe err", np.abs(fe - rfe).max(), "last", fe[-1],
                  "x_mean relL2", np.linalg.norm(xm - ref["x_mean"][:, :, 0]) / np.linalg.norm(ref["x_mean"]),
                  "x_cov rel", np.linalg.norm(xc - ref["x_cov"][:, :, :, 0]) / np.linalg.norm(ref["x_cov"]),
                  "theta err", np.abs(mth - ref["theta_mean"][-1, :, 0]).max(), "gamma", ga / gb, ref["gamma_shape"][-1, 0] / ref["gamma_rate"][-1, 0])
