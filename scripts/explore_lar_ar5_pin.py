"""Exploration (not shipped, not imported): does the reference's handling of the SINGULAR last-step message explain the
0.007 between the oracle's exact-limit AR(5) free energy (514.65389) and the reference pin (514.66086)?
The last x_T receives only the observation message (xi, W) = (c tau y, tau c c'), rank 1; the AR rules call
mean_cov(m_y) = cholinv(W) on it.  Hypothesis: the modified Cholesky behind cholinv replaces the zero pivots by `reg`."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import vmp
from scipy.special import digamma, gammaln

inv = np.linalg.inv


def run(y, p, tau, iters, reg, untelescoped=True):
    T = len(y)
    I = np.eye(p); c = np.zeros(p); c[0] = 1; ccT = np.outer(c, c)
    mth = np.zeros(p); Vth = np.eye(p); ga = gb = 1.0
    fes = []
    for _ in range(iters):
        Eg = ga / gb
        A = np.zeros((p, p)); A[0] = mth
        for i in range(1, p): A[i, i - 1] = 1
        V = np.zeros((p, p)); V[0, 0] = 1 / Eg
        gV = Eg * Vth
        # forward AR(:y) messages (mean, cov) and variable-side forward info
        Wf = np.zeros((T + 1, p, p)); xf = np.zeros((T + 1, p)); Wf[0] = I
        fm = np.zeros((T + 1, p)); fV = np.zeros((T + 1, p, p))
        for t in range(1, T + 1):
            Dinv = inv(Wf[t - 1] + gV); C = A @ Dinv
            my = C @ xf[t - 1]; Vy = C @ A.T + V
            fm[t], fV[t] = my, Vy
            Wy = inv(Vy)
            Wf[t] = Wy + tau * ccT; xf[t] = Wy @ my + tau * y[t - 1] * c
        # backward: message into AR_t's y interface
        Wb_true = tau * ccT; xb = tau * y[T - 1] * c
        Wb = Wb_true + reg * (I - ccT)              # what cholinv(cholinv(W)) leaves of the singular last-step message
        a = np.concatenate([[1.0], -mth])
        mu = np.zeros((T + 1, p + 1)); Su = np.zeros((T + 1, p + 1, p + 1))
        vm = np.zeros((T + 1, p)); vV = np.zeros((T + 1, p, p))
        Wb_var = Wb_true.copy(); xb_var = xb.copy()   # variable marginal uses the true product of messages
        for t in range(T, 0, -1):
            Wv = inv(fV[t]) + Wb_var; vV[t] = inv(Wv); vm[t] = vV[t] @ (inv(fV[t]) @ fm[t] + xb_var)
            L = Eg * np.outer(a, a); L[:p, :p] += Wb; L[1:, 1:] += Wf[t - 1] + gV
            eta = np.zeros(p + 1); eta[:p] += xb; eta[1:] += xf[t - 1]
            Su[t] = inv(L); mu[t] = Su[t] @ eta
            if t > 1:
                M = inv(I + Wb @ V); AM = A.T @ M
                Wx = AM @ Wb @ A + gV; xx = AM @ xb
                Wb = Wx + tau * ccT; xb = xx + tau * y[t - 2] * c
                Wb_var = Wb; xb_var = xb
            else:
                M = inv(I + Wb @ V); AM = A.T @ M
                W0 = AM @ Wb @ A + gV + I; x0 = AM @ xb
                vV[0] = inv(W0); vm[0] = vV[0] @ x0
        my1 = mu[1:, 0]; mx = mu[1:, 1:]; Vy1 = Su[1:, 0, 0]; Vy1x = Su[1:, 0, 1:]; Vx = Su[1:, 1:, 1:]
        Cx = Vx + np.einsum("ti,tj->tij", mx, mx); Lx = Vy1x + mx * my1[:, None]; Rx = Vy1 + my1 ** 2
        W = I + Eg * Cx.sum(0); Vth = inv(W); mth = Vth @ (Eg * Lx.sum(0))
        Bt = Rx - 2 * Lx @ mth + np.einsum("i,tij,j->t", mth, Cx, mth) + np.einsum("ij,tji->t", Vth, Cx)
        ga = 1 + 0.5 * T; gb = 1 + 0.5 * Bt.sum()
        Egn = ga / gb; Elog = digamma(ga) - np.log(gb)
        U_ar = 0.5 * (np.log(2 * np.pi) - Elog + Egn * Bt)
        H_u = 0.5 * ((p + 1) * (1 + np.log(2 * np.pi)) + np.linalg.slogdet(Su[1:])[1])
        H_x = 0.5 * (p * (1 + np.log(2 * np.pi)) + np.linalg.slogdet(vV[1:])[1])
        U_obs = 0.5 * (np.log(2 * np.pi) - np.log(tau) + tau * ((y - vm[1:, 0]) ** 2 + vV[1:, 0, 0]))
        U_x0 = 0.5 * (p * np.log(2 * np.pi) + np.trace(vV[0]) + vm[0] @ vm[0])
        kl_t = 0.5 * (np.trace(Vth) + mth @ mth - p - np.linalg.slogdet(Vth)[1])
        kl_g = (ga - 1) * digamma(ga) - gammaln(ga) + ga * (1 - gb) / gb + np.log(gb)
        fes.append(kl_t + kl_g + U_x0 + (U_ar - H_u).sum() + H_x[:-1].sum() + U_obs.sum())
    return np.array(fes)


if __name__ == "__main__":
    st, obs = vmp.latent_ar_reference_data()
    print("oracle exact:", vmp.latent_ar(obs[:, None], 5, 5.0, 15)["free_energy"][-1, 0], " pin 514.66086")
    for reg in (0.0, 1e-8, 0.1, 0.15, 0.5, 1.0, 5.0, 25.0):
        print("reg", reg, run(obs, 5, 5.0, 15, reg)[-1])
    print("AR1 reg=1:", run(obs, 1, 5.0, 15, 1.0)[-1], " pin 518.9182342")
