"""fp64 oracle for the Gamma-precision variational (VMP) rules and small fixed-data goldens.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

from . import rules as R


def gamma_aliases_golden(iterations=100, y=10.0, n=6):
    """Replays /root/reference/test/models/aliases/aliases_gamma_tests.jl:2-45:

        gamma[i] ~ Gamma(1, 1); x[i] ~ Normal(mu = 1, tau = gamma[i]), i = 1..6
        s ~ x[1] + ... + x[6];  y ~ Normal(mu = s, var = 1), y = 10
        constraints q(x, gamma) = q(x) q(gamma); init vague; 100 iterations.

    One iteration = sum-product over the Gaussian sub-graph (Normal(:out) VMP messages given
    E[gamma], Addition rules, product at every x[i]) followed by the Gamma update.
    Returns (mean(q(s)) per iteration, Bethe free energy per iteration).
    Goldens: mean 9.468846338832027, BFE 4.385584096993327 (test lines 43-44).
    """
    q_gamma = [(1.0, R.TINY)] * n            # vague(GammaShapeRate)
    means, fes = [], []
    for _ in range(iterations):
        # VMP message Normal(:out)(q_mu = PointMass(1), q_tau = q(gamma_i)): N(1, 1/E[gamma_i])
        px = [R.normal_meanprec_out_q_tau((1.0, 0.0), R.gamma_mean(g)) for g in q_gamma]
        # forward through the addition chain ((x1 + x2) + x3) + ... (rule +(:out))
        fwd = [px[0]]
        for i in range(1, n):
            fwd.append(R.addition_out(fwd[-1], px[i]))
        # backward: message into s from the observation, then +(:in1)/(:in2)
        back_s = (y, 1.0)                                   # NormalMeanVariance(:mu) from data
        q_s = R.prod_normal_mv(fwd[-1], back_s)
        qx = [None] * n
        back = back_s
        for i in range(n - 1, 0, -1):
            to_xi = R.addition_in2(back, fwd[i - 1])        # message to x[i]
            qx[i] = R.prod_normal_mv(px[i], to_xi)
            back = R.addition_in1(back, px[i])              # message to the partial sum
        qx[0] = R.prod_normal_mv(px[0], back)
        # Gamma update: prior Gamma(1,1) x NormalMeanPrecision(:tau)(q_out = q(x_i), q_mu = PointMass 1)
        q_gamma = [R.prod_gamma((1.0, 1.0), R.normal_meanprec_tau(q, (1.0, 0.0))) for q in qx]
        means.append(q_s[0])
        fes.append(_gamma_aliases_free_energy(qx, q_gamma, px, y))
    return np.array(means), np.array(fes)


def _gamma_aliases_free_energy(qx, q_gamma, px, y):
    """Variational free energy F[q] = E_q[log q - log p] of the model above under
    q(x) q(gamma); q(x) is the exact joint Gaussian given E[gamma] (BP on the addition tree),
    which is what the Bethe free energy (/root/reference/src/model/plugins/
    reactivemp_free_energy.jl:84-126, docs/src/manuals/variational/bethe-free-energy.md:45-52)
    evaluates to for this graph."""
    n = len(qx)
    Eg = np.array([R.gamma_mean(g) for g in q_gamma])
    Elog = np.array([R.gamma_mean_log(g) for g in q_gamma])
    # joint q(x): prior precision diag(p_i) with the p used for the BP sweep (px), obs y = 1'x + N(0,1)
    p = np.array([1.0 / v for (_, v) in px])
    W = np.diag(p) + np.ones((n, n))
    S = np.linalg.inv(W)
    m = S @ (p * 1.0 + y * np.ones(n))
    H_x = R.mvnormal_entropy(S)
    H_g = sum(R.gamma_entropy(g) for g in q_gamma)
    # -E log p(y | x)
    s_m = m.sum(); s_v = S.sum()
    U_y = 0.5 * (np.log(2 * np.pi) + (y - s_m) ** 2 + s_v)
    # -E log p(x_i | gamma_i), mu = 1
    U_x = sum(0.5 * (np.log(2 * np.pi) - Elog[i] + Eg[i] * ((m[i] - 1.0) ** 2 + S[i, i])) for i in range(n))
    # -E log p(gamma_i), Gamma(1,1): -log p = gamma
    U_g = Eg.sum()
    return U_y + U_x + U_g - H_x - H_g


def two_node_gaussian(prior_mean=3.0, prior_var=1.0, y=0.0, obs_var=1.0):
    """x ~ N(prior_mean, prior_var); y ~ N(x, obs_var) with y observed
    (/root/reference/test/models/models_tests.jl:226-256, 294-336).
    Returns (posterior mean, posterior var, BFE = -log evidence)."""
    m, v = R.prod_normal_mv((prior_mean, prior_var), (y, obs_var))
    s = prior_var + obs_var
    bfe = 0.5 * (np.log(2 * np.pi * s) + (y - prior_mean) ** 2 / s)
    return m, v, bfe


def lgssm_gamma_precision(y, A_scalar=1.0, prior=(0.0, 100.0), proc_var=1.0,
                          gamma_prior=(1.0, 1.0), iterations=10, init_Etau=1.0, return_free_energy=False):
    """Univariate smoother with an unknown SHARED observation precision tau ~ Gamma(a0, b0)
    (rules row 9): y[t] ~ N(x[t], 1/tau), x[t] ~ N(a x[t-1], v).
    Mean-field q(x) q(tau).  y[T, batch].  One VMP iteration = BP sweep given E[tau], then
    tau update with shapes/rates adding over t: Gamma(a0 + T/2, b0 + 1/2 sum E[(y - x)^2]).

    ``return_free_energy``: Bethe free energy after every iteration [iterations, batch], evaluated FROM ITS DEFINITION
    (/root/reference/src/model/plugins/reactivemp_free_energy.jl:84-126; q(x) = the exact Gaussian chain posterior of the
    sweep, dense T x T algebra): F = E_q[-log p(y, x, tau)] - H[q(x)] - H[q(tau)].  tests/test_oracle_goldens.py checks
    the closed form the CUDA kernel uses (filter evidence + T/2 (log tau_old - E log tau) + (E tau - tau_old)(b - b0) +
    KL(q(tau) || prior)) against it."""
    from scipy.special import digamma, gammaln
    from .lgssm import smooth_reference_schedule
    y = np.asarray(y, dtype=np.float64)
    T, batch = y.shape
    Etau = np.full(batch, init_Etau, dtype=np.float64)
    a0, b0 = gamma_prior
    fes, fes_closed = [], []
    # Gaussian chain prior over x[0..T-1]: precision J0 (tridiagonal), information vector h0
    J0 = np.zeros((T, T)); h0 = np.zeros(T)
    J0[0, 0] += 1.0 / prior[1]; h0[0] += prior[0] / prior[1]
    for t in range(1, T):
        J0[t, t] += 1.0 / proc_var; J0[t - 1, t - 1] += A_scalar ** 2 / proc_var
        J0[t, t - 1] -= A_scalar / proc_var; J0[t - 1, t] -= A_scalar / proc_var
    logdet_prior = -np.linalg.slogdet(J0)[1]                     # log det Sigma_prior
    mu_prior = np.linalg.solve(J0, h0)
    for _ in range(iterations):
        tau_old = Etau.copy()
        Q = (1.0 / Etau)[:, None, None]
        r = smooth_reference_schedule(
            y[:, None, :], np.array([[A_scalar]]), np.array([[1.0]]), np.array([[proc_var]]), Q,
            np.array([prior[0]]), np.array([[prior[1]]]))
        mx = r["mean"][:, 0, :]; vx = r["cov"][:, 0, 0, :]
        a = a0 + 0.5 * T
        b = b0 + 0.5 * ((y - mx) ** 2 + vx).sum(0)
        Etau = a / b
        if return_free_energy:
            Elog = digamma(a) - np.log(b)
            kl = (a - a0) * digamma(a) - gammaln(a) + gammaln(a0) + a0 * (np.log(b) - np.log(b0)) + a * (b0 - b) / b
            fes_closed.append(r["neg_log_evidence"] + 0.5 * T * (np.log(tau_old) - Elog) + (Etau - tau_old) * (b - b0) + kl)
            fe = np.zeros(batch)
            for c in range(batch):
                J = J0 + tau_old[c] * np.eye(T)                  # q(x) = N(mu, Sigma) under the E[tau] of the sweep
                Sig = np.linalg.inv(J)
                mu = Sig @ (h0 + tau_old[c] * y[:, c])
                dmu = mu - mu_prior
                E_prior = 0.5 * (T * np.log(2 * np.pi) + logdet_prior + np.trace(J0 @ Sig) + dmu @ J0 @ dmu)
                Rt = (y[:, c] - mu) ** 2 + np.diag(Sig)
                E_obs = 0.5 * (T * np.log(2 * np.pi) - T * Elog[c] + Etau[c] * Rt.sum())
                H_x = 0.5 * (T * np.log(2 * np.pi * np.e) + np.linalg.slogdet(Sig)[1])
                E_tau = -(a0 * np.log(b0) - gammaln(a0) + (a0 - 1) * Elog[c] - b0 * Etau[c])
                H_tau = a - np.log(b[c]) + gammaln(a) + (1 - a) * digamma(a)
                fe[c] = E_prior + E_obs + E_tau - H_x - H_tau
            fes.append(fe)
    out = dict(mean=mx, var=vx, shape=np.full(batch, a), rate=b, Etau=Etau)
    if return_free_energy:
        out["free_energy"] = np.stack(fes)
        out["free_energy_closed_form"] = np.stack(fes_closed)
    return out


def stream_vmp_gamma(y, iterations=4, w=1.0, init_x=(0.0, 1e3), init_tau=(1.0, 1.0), return_free_energy=False):
    """Streaming mean-field VMP of the reference's ``test_model1`` (one-step Kalman-like model with unknown
    observation precision, /root/reference/test/inference/inference_tests.jl:752-775):

        x_t_min ~ Normal(mean = x_t_min_mean, variance = x_t_min_var)     # datavars <- mean_var(q(x_t))
        tau     ~ Gamma(shape = tau_shape, rate = tau_rate)               # datavars <- shape / rate of q(tau)
        x_t     ~ Normal(mean = x_t_min, precision = w)                   # w = 1
        y       ~ Normal(mean = x_t, precision = tau)
        constraints = MeanField();  init q(x_t) = N(0, 1e3), q(tau) = Gamma(1, 1);  `iterations` per datum.

    Rules: NormalMeanPrecision(:out)(q_mu, q_tau) = N(E mu, 1/E tau), (:mu) symmetric,
    (:tau)(q_out, q_mu) = Gamma(3/2, 1/2 E(out - mu)^2), prod of Normals / Gammas.  Update order inside an iteration is
    not pinned by the reference (reactive); dependency order used here: q(x_t_min), q(x_t), q(tau).
    y[T, batch] -> out[T, 4, batch] = (m_x, v_x, shape, rate) after the last iteration of every datum, and (optional) the
    Bethe free energy [T, iterations, batch] (fully factorised q: F = sum_nodes U - sum_vars H)."""
    from scipy.special import digamma, gammaln
    y = np.asarray(y, dtype=np.float64)
    T, batch = y.shape
    mx = np.full(batch, init_x[0]); vx = np.full(batch, init_x[1])
    a = np.full(batch, init_tau[0]); b = np.full(batch, init_tau[1])
    out = np.zeros((T, 4, batch))
    fe = np.zeros((T, iterations, batch)) if return_free_energy else None
    for t in range(T):
        mp, vp, ap, bp = mx.copy(), vx.copy(), a.copy(), b.copy()       # autoupdates: priors of this datum
        for it in range(iterations):
            mmin, vmin = R.prod_normal_mv((mp, vp), (mx, 1.0 / w))      # q(x_t_min) = prior x NormalMeanPrecision(:mu)(q_out)
            Etau = a / b
            mx, vx = R.prod_normal_mv((mmin, 1.0 / w), (y[t], 1.0 / Etau))
            ga, gb = R.normal_meanprec_tau((y[t], 0.0), (mx, vx))
            a, b = R.prod_gamma((ap, bp), (ga, gb))
            if return_free_energy:
                Etau = a / b; Elog = digamma(a) - np.log(b)
                U1 = 0.5 * np.log(2 * np.pi * vp) + ((mmin - mp) ** 2 + vmin) / (2 * vp)
                U2 = -ap * np.log(bp) + gammaln(ap) - (ap - 1.0) * Elog + bp * Etau
                U3 = 0.5 * np.log(2 * np.pi) - 0.5 * np.log(w) + 0.5 * w * ((mx - mmin) ** 2 + vx + vmin)
                U4 = 0.5 * np.log(2 * np.pi) - 0.5 * Elog + 0.5 * Etau * ((y[t] - mx) ** 2 + vx)
                H = R.normal_entropy(vmin) + R.normal_entropy(vx) + R.gamma_entropy((a, b))
                fe[t, it] = U1 + U2 + U3 + U4 - H
        out[t, 0], out[t, 1], out[t, 2], out[t, 3] = mx, vx, a, b
    return (out, fe) if return_free_energy else out


def mv_iid_wishart(y, iterations=10, mu0=None, Lambda0=None, nu0=None, inv_scale0=None, init_E_P=None):
    """Mean-field VMP of /root/reference/test/models/iid/mv_iid_precision_tests.jl:10-41, message by message:

        m ~ MvNormal(mu = 0, Lambda = 100 I);  P ~ Wishart(d + 1, I);  y[i] ~ MvNormal(mu = m, Lambda = P);  q(m) q(P)

    y[N, d, batch] -> dict(m_mean[d, batch], m_cov[d, d, batch], df[batch], inv_scale[d, d, batch]).
    Per iteration: q(m) = prior x prod_i MvNormalMeanPrecision(:mu)(q_out = y_i, q_Lambda) (precision E[P] each);
    q(P) = prior x prod_i MvNormalMeanPrecision(:Lambda)(q_out = y_i, q_mu) (rules.mvnormal_meanprec_lambda / prod_wishart).
    The initial q(P) enters through its mean only (``vague(Wishart, d)`` in the reference: df = d, scale 1e12 I)."""
    y = np.asarray(y, dtype=np.float64)
    N, d, batch = y.shape
    mu0 = np.zeros(d) if mu0 is None else np.asarray(mu0, np.float64)
    Lambda0 = 100.0 * np.eye(d) if Lambda0 is None else np.asarray(Lambda0, np.float64)
    nu0 = d + 1.0 if nu0 is None else float(nu0)
    inv_scale0 = np.eye(d) if inv_scale0 is None else np.asarray(inv_scale0, np.float64)
    EP = np.broadcast_to(d * 1e12 * np.eye(d) if init_E_P is None else np.asarray(init_E_P, np.float64), (batch, d, d)).copy()
    yb = np.moveaxis(y, 2, 0)                                  # [batch, N, d]
    for _ in range(iterations):
        Lm = Lambda0 + N * EP
        Vm = np.linalg.inv(Lm)
        xi = Lambda0 @ mu0 + np.einsum("bij,bj->bi", EP, yb.sum(axis=1))
        m = np.einsum("bij,bj->bi", Vm, xi)
        w = (np.full(batch, nu0), np.broadcast_to(inv_scale0, (batch, d, d)).copy())
        # fold the N Wishart messages (left to right, as the reference's product fold does)
        df_acc, is_acc = w
        msg = R.mvnormal_meanprec_lambda((yb, np.zeros((batch, N, d, d))), (m[:, None, :], Vm[:, None, :, :]))
        for i in range(N):
            df_acc, is_acc = R.prod_wishart((df_acc, is_acc), (msg[0][:, i], msg[1][:, i]))
        EP = R.wishart_mean((df_acc, is_acc))
    return dict(m_mean=m.T.copy(), m_cov=np.moveaxis(Vm, 0, 2).copy(), df=df_acc, inv_scale=np.moveaxis(is_acc, 0, 2).copy(),
                E_P=np.moveaxis(EP, 0, 2).copy())


def ar_regression(series, order, iterations=15, gamma_prior=(1.0, 1.0), theta_prior_precision=1.0, init_gamma=(1.0, 1.0)):
    """Mean-field VMP of the reference's autoregressive test model, message by message
    (/root/reference/test/models/autoregressive/ar_tests.jl:7-36):

        gamma ~ Gamma(shape = 1, rate = 1);  theta ~ MvNormal(mean = 0, precision = I)
        y[i] ~ Normal(mean = dot(x[i], theta), precision = gamma),   x[i] = lags of the series (ar_ssm_data, :7-15)
        q(gamma, theta) = q(gamma) q(theta);  init q(gamma) = GammaShapeRate(1, 1);  `iterations` sweeps

    series[N, batch] -> dict(theta_mean[order, batch], theta_cov[order, order, batch], gamma_shape, gamma_rate[batch],
    free_energy[iterations, batch]).  Per sweep: every Normal node sends, through the `dot` node with the PointMass x[i], the
    backward message (xi, W) = (x_i E[gamma] y_i, E[gamma] x_i x_i') to theta (rules: NormalMeanPrecision(:mu)(q_out, q_tau),
    dot(:in2)); q(theta) = prior x product of them; then NormalMeanPrecision(:tau)(q_out = y_i, q_mu = q(x_i'theta)) =
    Gamma(3/2, 1/2 [(y_i - x_i'm)^2 + x_i' V x_i]) and q(gamma) = prior x product.  Bethe free energy of the fully factorised q:
    E[-log p(y | theta, gamma)] + KL(q(theta) || p(theta)) + KL(q(gamma) || p(gamma))."""
    from scipy.special import digamma, gammaln
    s = np.asarray(series, dtype=np.float64)
    N, batch = s.shape
    p = order
    n = N - p
    # ar_ssm_data: inputs[k] = (s[k+p-1], ..., s[k]) reversed window, outputs[k] = s[k+p]
    X = np.stack([s[p - 1 - j: N - 1 - j] for j in range(p)], axis=1)          # [n, p, batch]
    Y = s[p:]                                                                  # [n, batch]
    a0, b0 = gamma_prior
    ga = np.full(batch, init_gamma[0]); gb = np.full(batch, init_gamma[1])
    fes = []
    for _ in range(iterations):
        Eg = ga / gb
        W = theta_prior_precision * np.eye(p)[:, :, None] + Eg * np.einsum("nib,njb->ijb", X, X)
        xi = Eg * np.einsum("nib,nb->ib", X, Y)
        V = np.linalg.inv(np.moveaxis(W, 2, 0))                                # [batch, p, p]
        m = np.einsum("bij,jb->bi", V, xi)                                     # [batch, p]
        pred = np.einsum("nib,bi->nb", X, m)
        xVx = np.einsum("nib,bij,njb->nb", X, V, X)
        res = ((Y - pred) ** 2 + xVx).sum(0)
        ga = a0 + 0.5 * n + 0.0 * ga                                           # a0 + sum (3/2 - 1)
        gb = b0 + 0.5 * res
        Elog = digamma(ga) - np.log(gb); Egn = ga / gb
        like = 0.5 * n * (np.log(2 * np.pi) - Elog) + 0.5 * Egn * res
        klt = 0.5 * (theta_prior_precision * (np.trace(V, axis1=1, axis2=2) + (m ** 2).sum(1)) - p - p * np.log(theta_prior_precision)
                     - np.linalg.slogdet(V)[1])
        klg = (ga - a0) * digamma(ga) - gammaln(ga) + gammaln(a0) + a0 * (np.log(gb) - np.log(b0)) + ga * (b0 - gb) / gb
        fes.append(like + klt + klg)
    return dict(theta_mean=m.T.copy(), theta_cov=np.moveaxis(V, 0, 2).copy(), gamma_shape=ga, gamma_rate=gb, free_energy=np.stack(fes))


def latent_ar_reference_data(n=500, theta=None, gamma=5.0, tau=5.0, seed=123):
    """/root/reference/test/models/autoregressive/lar_tests.jl:128-157 on the regenerated StableRNG stream:
    states[1] = randn(rng, order); states[i] = [Normal(theta'states[i-1], 1/sqrt(gamma)); states[i-1][1:end-1]];
    observations[i] ~ Normal(states[i][1], tau_std) with tau_std = sqrt(inv(gamma)) (sic, :137); the first 3*order
    entries are dropped.  Returns (states[n, order], observations[n])."""
    from .julia_rng import StableRNG
    if theta is None:
        theta = [0.10699399235785655, -0.5237303489793305, 0.3068897071844715, -0.17232255282458891, 0.13323964347539288]
    theta = np.asarray(theta, dtype=np.float64)
    p = len(theta)
    rng = StableRNG(seed)
    g_std = np.sqrt(1.0 / gamma)
    t_std = np.sqrt(1.0 / gamma)
    N = n + 3 * p
    states = np.zeros((N, p)); obs = np.zeros(N)
    states[0] = rng.randn_vec(p)
    for i in range(1, N):
        head = float(theta @ states[i - 1]) + g_std * rng.randn()
        states[i] = np.concatenate([[head], states[i - 1][:-1]])
        obs[i] = states[i][0] + t_std * rng.randn()
    return states[3 * p:], obs[3 * p:]


def latent_ar(y, order, tau, iterations=15, gamma_prior=(1.0, 1.0), init_gamma=(1.0, 1.0), schedule="x_theta_gamma",
              literal_huge=None, theta_prior_precision=1.0, x0_prior_precision=1.0, init_theta_precision=1.0):
    """Latent autoregressive model (/root/reference/test/models/autoregressive/lar_tests.jl:52-122):

        gamma ~ Gamma(1, 1); theta ~ N(0, I); x0 ~ N(0, I)
        x[t] ~ AR(x[t-1], theta, gamma)  (ARMeta(variate, order, ARsafe())),   y[t] ~ Normal(c'x[t], precision = tau), c = e1
        q(x, x0, gamma, theta) = q(x, x0) q(gamma) q(theta); init q(gamma) = Gamma(1, 1), q(theta) = N(0, I)

    y[T, batch] -> dict(x_mean[T, order, batch], x_cov[T, order, order, batch], theta_mean/cov per iteration, gamma per
    iteration, free_energy[iterations, batch]).  The AR node (ReactiveMP `AR`, rules restated from the published
    algorithm: the reference repo only holds the call sites and the free-energy pins :170, :201) is
        f(y, x, theta, gamma) = N(y1 | theta'x, 1/gamma) * prod_{i>1} delta(y_i - x_{i-1});
    ReactiveMP regularises the deltas with precision `huge` = 1e12 (`ARPrecisionMatrix`); this restatement works in
    the exact limit on u = (y1, x) (order + 1 coordinates; y = u[:order], x = u[1:]) and `literal_huge` switches to the
    regularised 2*order-dimensional form for the cross-check.  Rules:
        AR(:y)(m_x, q_theta, q_gamma):  D = W_x + E[gamma] V_theta, my = A D^-1 xi_x, Vy = A D^-1 A' + diag(1/E[gamma], 0...)
        AR(:x)(m_y, q_theta, q_gamma):  W = A'(Vy + V)^-1 A + E[gamma] V_theta, xi = A'(Vy + V)^-1 my   (A = companion(E theta))
        marginal q(y, x):               precision S_y'W_y S_y + S_x'(W_x + E[gamma] V_theta)S_x + E[gamma] a a', a = (1, -E theta)
        AR(:theta)(q_yx, q_gamma):      (xi, W) = E[gamma] (V_y1x + m_x m_y1,  V_x + m_x m_x')
        AR(:gamma)(q_yx, q_theta):      Gamma(3/2, B/2), B = E[(y1 - theta'x)^2]
    Bethe free energy: the AR node's average energy carries the reference's entropy correction (the degenerate
    coordinates of q(y, x) are dropped: H[q(y, x)] -> H[q(y1, x)]); deterministic `dot` / `*` nodes contribute -H[q(x_t)];
    summed over the graph:
        F = KL(q(theta)||p) + KL(q(gamma)||p) + E[-log p(x0)] + sum_t (U_AR,t - H[q(y1, x)_t]) + sum_{t<T} H[q(x_t)] + sum_t U_obs,t."""
    from scipy.special import digamma, gammaln
    y = np.asarray(y, dtype=np.float64)
    T, Bn = y.shape
    p = order
    I = np.eye(p)
    c = np.zeros(p); c[0] = 1.0
    ccT = np.outer(c, c)
    a0, b0 = gamma_prior
    mth = np.zeros((Bn, p)); Vth = np.tile(I / init_theta_precision, (Bn, 1, 1))
    ga = np.full(Bn, init_gamma[0]); gb = np.full(Bn, init_gamma[1])
    w0, p0 = theta_prior_precision, x0_prior_precision
    hist = dict(theta_mean=[], theta_cov=[], gamma_shape=[], gamma_rate=[], free_energy=[])
    inv = np.linalg.inv
    for _ in range(iterations):
        Eg = ga / gb
        A = np.zeros((Bn, p, p)); A[:, 0, :] = mth
        for i in range(1, p):
            A[:, i, i - 1] = 1.0
        AT = np.swapaxes(A, 1, 2)
        V = np.zeros((Bn, p, p)); V[:, 0, 0] = 1.0 / Eg
        gV = Eg[:, None, None] * Vth
        # ---- forward messages (xi, W) on x_0 .. x_T, observation included
        Wf = np.zeros((T + 1, Bn, p, p)); xf = np.zeros((T + 1, Bn, p))
        Wf[0] = p0 * I
        for t in range(1, T + 1):
            Dinv = inv(Wf[t - 1] + gV)
            C = A @ Dinv
            my = np.einsum("bij,bj->bi", C, xf[t - 1])
            Vy = C @ AT + V
            Wy = inv(Vy)
            Wf[t] = Wy + tau * ccT
            xf[t] = np.einsum("bij,bj->bi", Wy, my) + tau * y[t - 1][:, None] * c
        # ---- backward messages into AR_t's y interface, marginals q(u_t), u = (y1, x)
        Wb = np.tile(tau * ccT, (Bn, 1, 1)); xb = tau * y[T - 1][:, None] * c
        mu = np.zeros((T + 1, Bn, p + 1)); Su = np.zeros((T + 1, Bn, p + 1, p + 1))
        a = np.concatenate([np.ones((Bn, 1)), -mth], axis=1)
        for t in range(T, 0, -1):
            if literal_huge is None:
                L = Eg[:, None, None] * np.einsum("bi,bj->bij", a, a)
                L[:, :p, :p] += Wb
                L[:, 1:, 1:] += Wf[t - 1] + gV
                eta = np.zeros((Bn, p + 1)); eta[:, :p] += xb; eta[:, 1:] += xf[t - 1]
                S = inv(L)
                mu[t] = np.einsum("bij,bj->bi", S, eta); Su[t] = S
            else:
                mW = np.zeros((Bn, p, p)); mW[:, 0, 0] = Eg
                for i in range(1, p):
                    mW[:, i, i] = literal_huge
                W = np.zeros((Bn, 2 * p, 2 * p))
                W[:, :p, :p] = Wb + mW
                W[:, :p, p:] = -mW @ A
                W[:, p:, :p] = -AT @ mW
                W[:, p:, p:] = Wf[t - 1] + gV + AT @ mW @ A
                eta = np.concatenate([xb, xf[t - 1]], axis=1)
                S2 = inv(W); m2 = np.einsum("bij,bj->bi", S2, eta)
                idx = [0] + list(range(p, 2 * p))
                mu[t] = m2[:, idx]; Su[t] = S2[:, idx][:, :, idx]
            if t > 1:
                M = inv(I + Wb @ V)
                AM = AT @ M
                Wx = AM @ Wb @ A + gV
                xx = np.einsum("bij,bj->bi", AM, xb)
                Wb = Wx + tau * ccT
                xb = xx + tau * y[t - 2][:, None] * c
        my1 = mu[1:, :, 0]; mx = mu[1:, :, 1:]
        Vy1 = Su[1:, :, 0, 0]; Vy1x = Su[1:, :, 0, 1:]; Vx = Su[1:, :, 1:, 1:]
        Cx = Vx + np.einsum("tbi,tbj->tbij", mx, mx)
        Lx = Vy1x + mx * my1[:, :, None]
        Rx = Vy1 + my1 ** 2

        def theta_update(Eg_):
            W = w0 * I + Eg_[:, None, None] * Cx.sum(0)
            xi = Eg_[:, None] * Lx.sum(0)
            Vn = inv(W)
            return np.einsum("bij,bj->bi", Vn, xi), Vn

        def gamma_update(mth_, Vth_):
            Bt = (Rx - 2 * np.einsum("bi,tbi->tb", mth_, Lx) + np.einsum("bi,tbij,bj->tb", mth_, Cx, mth_)
                  + np.einsum("bij,tbji->tb", Vth_, Cx))
            return np.full(Bn, a0 + 0.5 * T), b0 + 0.5 * Bt.sum(0)

        if schedule == "x_theta_gamma":
            mth, Vth = theta_update(Eg)
            ga, gb = gamma_update(mth, Vth)
        elif schedule == "x_gamma_theta":
            ga, gb = gamma_update(mth, Vth)
            mth, Vth = theta_update(ga / gb)
        else:                                            # "jacobi": both from the previous iteration's marginals
            mth_n, Vth_n = theta_update(Eg)
            ga, gb = gamma_update(mth, Vth)
            mth, Vth = mth_n, Vth_n
        # ---- Bethe free energy with the new q(theta), q(gamma) and the q(u_t) just computed
        Egn = ga / gb; Elog = digamma(ga) - np.log(gb)
        Bt = (Rx - 2 * np.einsum("bi,tbi->tb", mth, Lx) + np.einsum("bi,tbij,bj->tb", mth, Cx, mth)
              + np.einsum("bij,tbji->tb", Vth, Cx))
        U_ar = 0.5 * (np.log(2 * np.pi) - Elog[None] + Egn[None] * Bt)
        H_u = 0.5 * ((p + 1) * (1 + np.log(2 * np.pi)) + np.linalg.slogdet(Su[1:])[1])
        # q(x_t) = marginal of the y block of q(y, x)_t: y = u[:p]
        mxt = mu[1:, :, :p]; Vxt = Su[1:, :, :p, :p]
        H_x = 0.5 * (p * (1 + np.log(2 * np.pi)) + np.linalg.slogdet(Vxt)[1])
        U_obs = 0.5 * (np.log(2 * np.pi) - np.log(tau) + tau * ((y - mxt[:, :, 0]) ** 2 + Vxt[:, :, 0, 0]))
        m0 = mu[1, :, 1:]; V0 = Su[1, :, 1:, 1:]
        U_x0 = 0.5 * (p * np.log(2 * np.pi) - p * np.log(p0) + p0 * (np.trace(V0, axis1=1, axis2=2) + (m0 ** 2).sum(1)))
        kl_t = 0.5 * (w0 * (np.trace(Vth, axis1=1, axis2=2) + (mth ** 2).sum(1)) - p - p * np.log(w0) - np.linalg.slogdet(Vth)[1])
        kl_g = (ga - a0) * digamma(ga) - gammaln(ga) + gammaln(a0) + a0 * (np.log(gb) - np.log(b0)) + ga * (b0 - gb) / gb
        fe = kl_t + kl_g + U_x0 + (U_ar - H_u).sum(0) + H_x[:-1].sum(0) + U_obs.sum(0)
        hist["theta_mean"].append(mth.T.copy()); hist["theta_cov"].append(np.moveaxis(Vth, 0, 2).copy())
        hist["gamma_shape"].append(ga.copy()); hist["gamma_rate"].append(gb.copy()); hist["free_energy"].append(fe)
    out = {k: np.stack(v) for k, v in hist.items()}
    out["x_mean"] = np.moveaxis(mxt, 1, 2).copy()                    # [T, order, batch]
    out["x_cov"] = np.moveaxis(Vxt, 1, 3).copy()                     # [T, order, order, batch]
    return out
