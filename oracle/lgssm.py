"""fp64 oracle for the linear-Gaussian state-space hot path (smoothing and filtering).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Model (reference: /root/reference/benchmarks/Linear Multivariate Gaussian State Space Model
Benchmark.ipynb:95-105, test/models/statespace/mlgssm_test.jl:8-17):

    x[1] ~ MvNormal(m0, S0);  x[t] ~ MvNormal(A x[t-1], P);  y[t] ~ MvNormal(B x[t], Q)

``smooth_reference_schedule`` replays the *message schedule* the reference executes for
``infer(model = linear_gaussian_ssm_smoothing(...), data = (y = ...,))`` -- rule by rule, in
the parametrisation each rule produces -- rather
than a textbook smoother, so that intermediate messages can be compared too.
``kalman_rts`` is the independent textbook cross-check.

Array layouts follow the C ABI (batch innermost):
    y[T, m, batch]   mean[T, d, batch]   cov[T, d, d, batch]   mask[T, batch] (1 = observed)
Model matrices are shared ([d, d]) or per chain ([batch, d, d]).
"""
from __future__ import annotations

import numpy as np

from . import rules as R


def _bc(M, batch, shape):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == len(shape):
        return np.broadcast_to(M, (batch,) + tuple(shape)).copy()
    assert M.shape == (batch,) + tuple(shape), (M.shape, batch, shape)
    return M


def _unpack(y, A, B, P, Q, m0, S0, mask, u=None):
    y = np.asarray(y, dtype=np.float64)
    T, m, batch = y.shape
    d = np.asarray(A).shape[-1]
    yb = np.ascontiguousarray(np.transpose(y, (0, 2, 1)))        # [T, batch, m]
    A = _bc(A, batch, (d, d)); B = _bc(B, batch, (m, d))
    P = _bc(P, batch, (d, d)); Q = _bc(Q, batch, (m, m))
    m0 = _bc(m0, batch, (d,)); S0 = _bc(S0, batch, (d, d))
    _unpack.u = _bc(np.zeros(d) if u is None else u, batch, (d,))
    if mask is None:
        mk = np.ones((T, batch), dtype=bool)
    else:
        mk = np.asarray(mask).astype(bool)
    return yb, A, B, P, Q, m0, S0, mk, T, m, d, batch


def _pack(mu, S):
    """[T, batch, d] / [T, batch, d, d] -> ABI layout."""
    return (np.ascontiguousarray(np.transpose(mu, (0, 2, 1))),
            np.ascontiguousarray(np.transpose(S, (0, 2, 3, 1))))


def observation_message(yt, mk_t, B, Q):
    """Rules #3 (MvNormalMeanCovariance(:mu) from data) then #4 (*(:in)) for one time step.
    Missing datum => no message, i.e. (xi, W) = (0, 0)
    (semantics: /root/reference/docs/src/manuals/inference/static.md:98-125)."""
    mu_o, S_o = R.mvnormal_meancov_mean_from_data(yt, Q)
    xi_o, W_o = R.meancov_to_wmp(mu_o, S_o)
    xi, W = R.multiplication_in((xi_o, W_o), B)
    keep = mk_t[:, None]
    return xi * keep, W * keep[..., None]


def smooth_reference_schedule(y, A, B, P, Q, m0, S0, mask=None, return_messages=False, u=None,
                              transition_first=False):
    """Forward/backward sum-product sweep exactly as the reference schedules it.

    Returns dict(mean, cov, filt_mean, filt_cov, neg_log_evidence[batch]) in ABI layout.
    Message count: 6 rule calls per (chain, step) (SURVEY.md section 8a accounting).

    ``u``: constant transition offset, ``x[t] ~ MvNormal(A x[t-1] + u, P)`` -- in the reference graph
    an Addition node with a PointMass operand between ``*`` and the MvNormal node (pure mean shift,
    rules +(:out) forward and +(:in1) backward; ``x[i] ~ x_prev + c`` at
    test/models/statespace/ulgssm_tests.jl:12).  ``transition_first``: the prior sits on the state
    before x[1] (``x_prior ~ MvNormal(mean(x0), cov(x0)); x[1] ~ MvNormal(A x_prior, Q)``,
    test/models/statespace/mlgssm_test.jl:8-17).
    """
    yb, A, B, P, Q, m0, S0, mk, T, m, d, batch = _unpack(y, A, B, P, Q, m0, S0, mask, u)
    u = _unpack.u
    shift = lambda msg, sgn: (msg[0] + sgn * u, msg[1])      # +(:out) / +(:in1) with a PointMass operand

    fwd_mu = np.zeros((T, batch, d)); fwd_S = np.zeros((T, batch, d, d))
    obs_xi = np.zeros((T, batch, d)); obs_W = np.zeros((T, batch, d, d))
    fil_mu = np.zeros((T, batch, d)); fil_S = np.zeros((T, batch, d, d))
    nle = np.zeros(batch)

    # ---- forward: prior, then (#3,#4) observation, product, (#1) A*x, (#2) +P
    f_mu, f_S = R.mvnormal_meancov_out((m0, np.zeros_like(S0)), S0)      # prior, rule #2 with PointMass mean
    if transition_first:
        f_mu, f_S = R.mvnormal_meancov_out(shift(R.multiplication_out(A, (f_mu, f_S)), +1.0), P)
    for t in range(T):
        fwd_mu[t], fwd_S[t] = f_mu, f_S
        obs_xi[t], obs_W[t] = observation_message(yb[t], mk[t], B, Q)
        # log-evidence increment (innovation form == Bethe free energy on a tree)
        e = yb[t] - R.mv(B, f_mu)
        Sinn = B @ f_S @ np.swapaxes(B, -1, -2) + Q
        Sinv = R.cholinv(Sinn)
        inc = 0.5 * (m * np.log(2 * np.pi) + np.linalg.slogdet(Sinn)[1]
                     + np.einsum("bi,bij,bj->b", e, Sinv, e))
        nle += np.where(mk[t], inc, 0.0)
        # filtered = prod(fwd, obs) in (xi, W); back to (mu, Sigma) for rule #1
        xi_f, W_f = R.prod_gaussian_wmp(R.meancov_to_wmp(f_mu, f_S), (obs_xi[t], obs_W[t]))
        fil_mu[t], fil_S[t] = R.wmp_to_meancov(xi_f, W_f)
        f_mu, f_S = R.mvnormal_meancov_out(shift(R.multiplication_out(A, (fil_mu[t], fil_S[t])), +1.0), P)

    # ---- backward: bwd_T = none; out = prod(obs, bwd) -> (#3') +P -> (#4) A' . A
    post_mu = np.zeros((T, batch, d)); post_S = np.zeros((T, batch, d, d))
    bwd_xi_all = np.zeros((T, batch, d)); bwd_W_all = np.zeros((T, batch, d, d))
    b_xi = np.zeros((batch, d)); b_W = np.zeros((batch, d, d))
    for t in range(T - 1, -1, -1):
        bwd_xi_all[t], bwd_W_all[t] = b_xi, b_W
        post_mu[t], post_S[t] = R.marginal_from_messages(
            [R.meancov_to_wmp(fwd_mu[t], fwd_S[t]), (obs_xi[t], obs_W[t]), (b_xi, b_W)])
        if t == 0:
            break
        o_xi, o_W = R.prod_gaussian_wmp((obs_xi[t], obs_W[t]), (b_xi, b_W))
        # The reference converts (xi, W) -> (mu, Sigma) here (cholinv), adds P (rule #3'), converts
        # back (cholinv) and applies rule #4.  That needs W to be invertible.  When it is not --
        # no information from the future at all (trailing missing data), or m < d so that
        # B' Q^-1 B is rank deficient -- the reference falls into FastCholesky's
        # PositiveFactorizations repair (SURVEY.md 8c "not a usable golden"), which is not a
        # defined result.  The oracle then takes the analytic value of the same two rules,
        # (W^-1 + P)^-1 = (I + W P)^-1 W, which coincides with the reference form whenever W is SPD.
        ev = np.linalg.eigvalsh(R.sym(o_W))
        spd = ev[:, 0] > 1e-10 * np.maximum(ev[:, -1], 1e-300)
        o_W_safe = np.where(spd[:, None, None], o_W, np.eye(d))
        o_mu, o_S = R.wmp_to_meancov(o_xi, o_W_safe)
        n_mu, n_S = R.mvnormal_meancov_mean((o_mu, o_S), P)
        n_xi, n_W = R.meancov_to_wmp(n_mu, n_S)
        IWP = np.eye(d) + o_W @ P
        s_W = R.sym(np.linalg.solve(IWP, o_W))
        s_xi = np.linalg.solve(IWP, o_xi[..., None])[..., 0]
        n_xi = np.where(spd[:, None], n_xi, s_xi)
        n_W = np.where(spd[:, None, None], n_W, s_W)
        n_xi = n_xi - R.mv(n_W, u)                     # +(:in1): mean shifted by -u, in (xi, W) form
        b_xi, b_W = R.multiplication_in((n_xi, n_W), A)

    mean, cov = _pack(post_mu, post_S)
    fmean, fcov = _pack(fil_mu, fil_S)
    out = dict(mean=mean, cov=cov, filt_mean=fmean, filt_cov=fcov, neg_log_evidence=nle)
    if return_messages:
        out.update(fwd_mean=fwd_mu, fwd_cov=fwd_S, obs_xi=obs_xi, obs_W=obs_W,
                   bwd_xi=bwd_xi_all, bwd_W=bwd_W_all)
    return out


def filter_reference_schedule(y, A, B, P, Q, m0, S0, mask=None, u=None, transition_first=False):
    """Forward half only: what ``rxinfer_inference_filtering`` (ipynb:199-216) produces through
    ``@autoupdates x_min_t_mean, x_min_t_cov = mean_cov(q(x_t))``
    (/root/reference/src/inference/autoupdates.jl:614-659).

    NOTE the streaming model (ipynb:110-113) places the transition *before* the first datum:
    x_min_t ~ prior; x_t ~ N(A x_min_t, P); y_t ~ N(B x_t, Q).  ``transition_first=True`` in
    ``filter_streaming`` reproduces that; this function is the forward half of the smoothing
    graph (prior sits on x[1])."""
    r = smooth_reference_schedule(y, A, B, P, Q, m0, S0, mask, u=u, transition_first=transition_first)
    return dict(mean=r["filt_mean"], cov=r["filt_cov"], neg_log_evidence=r["neg_log_evidence"])


def filter_streaming(y, A, B, P, Q, m0, S0, mask=None, u=None):
    """Streaming filter as the notebook runs it: the prior (initialised to q(x_t) = N(m0, S0))
    is pushed through the transition before every datum, including the first."""
    yb, A, B, P, Q, m0, S0, mk, T, m, d, batch = _unpack(y, A, B, P, Q, m0, S0, mask, u)
    u = _unpack.u
    mu, S = m0, S0
    out_mu = np.zeros((T, batch, d)); out_S = np.zeros((T, batch, d, d))
    for t in range(T):
        p_mu, p_S = R.multiplication_out(A, (mu, S))
        p_mu, p_S = R.mvnormal_meancov_out((p_mu + u, p_S), P)
        o_xi, o_W = observation_message(yb[t], mk[t], B, Q)
        xi, W = R.prod_gaussian_wmp(R.meancov_to_wmp(p_mu, p_S), (o_xi, o_W))
        mu, S = R.wmp_to_meancov(xi, W)
        out_mu[t], out_S[t] = mu, S
    mean, cov = _pack(out_mu, out_S)
    return dict(mean=mean, cov=cov)


def kalman_rts(y, A, B, P, Q, m0, S0, mask=None, u=None, transition_first=False):
    """Textbook Kalman filter + Rauch-Tung-Striebel smoother (independent cross-check)."""
    yb, A, B, P, Q, m0, S0, mk, T, m, d, batch = _unpack(y, A, B, P, Q, m0, S0, mask, u)
    u = _unpack.u
    At = np.swapaxes(A, -1, -2); Bt = np.swapaxes(B, -1, -2)
    pm = np.zeros((T, batch, d)); pS = np.zeros((T, batch, d, d))
    fm = np.zeros((T, batch, d)); fS = np.zeros((T, batch, d, d))
    nle = np.zeros(batch)
    mu, S = m0, S0
    if transition_first:
        mu, S = R.mv(A, mu) + u, A @ S @ At + P
    for t in range(T):
        if t > 0:
            mu, S = R.mv(A, fm[t - 1]) + u, A @ fS[t - 1] @ At + P
        pm[t], pS[t] = mu, S
        Sinn = B @ S @ Bt + Q
        K = S @ Bt @ np.linalg.inv(Sinn)
        e = yb[t] - R.mv(B, mu)
        keep = mk[t]
        mu_u = mu + R.mv(K, e)
        S_u = S - K @ Sinn @ np.swapaxes(K, -1, -2)
        fm[t] = np.where(keep[:, None], mu_u, mu)
        fS[t] = np.where(keep[:, None, None], S_u, S)
        inc = 0.5 * (m * np.log(2 * np.pi) + np.linalg.slogdet(Sinn)[1]
                     + np.einsum("bi,bij,bj->b", e, np.linalg.inv(Sinn), e))
        nle += np.where(keep, inc, 0.0)
    sm = fm.copy(); sS = fS.copy()
    for t in range(T - 2, -1, -1):
        G = fS[t] @ At @ np.linalg.inv(pS[t + 1])
        sm[t] = fm[t] + R.mv(G, sm[t + 1] - pm[t + 1])
        sS[t] = fS[t] + G @ (sS[t + 1] - pS[t + 1]) @ np.swapaxes(G, -1, -2)
    mean, cov = _pack(sm, sS)
    fmean, fcov = _pack(fm, fS)
    return dict(mean=mean, cov=cov, filt_mean=fmean, filt_cov=fcov, neg_log_evidence=nle)


# --------------------------------------------------------------------------- synthetic data
def notebook_model(d=4):
    """The notebook's model (ipynb:157-162, 98) lifted to d = 2 or 4 as SURVEY.md section 8d
    config 1 prescribes: A = blkdiag(R(pi/15), R(pi/35)), B = diag(1.3, .7, 1.3, .7),
    P = 0.05 I, Q = 10 I, prior N(0, 100 I)."""
    def rot(th):
        return np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    if d == 2:
        A = rot(np.pi / 15); B = np.diag([1.3, 0.7])
    elif d == 4:
        A = np.zeros((4, 4)); A[:2, :2] = rot(np.pi / 15); A[2:, 2:] = rot(np.pi / 35)
        B = np.diag([1.3, 0.7, 1.3, 0.7])
    else:
        raise ValueError(d)
    return dict(A=A, B=B, P=0.05 * np.eye(d), Q=10.0 * np.eye(d),
                m0=np.zeros(d), S0=100.0 * np.eye(d))


def dense_model(d=64, seed=64):
    """A = 0.99 * Orth (Q-factor of a default_rng(seed) Gaussian
    d x d), B = I, P = 0.05 I, Q = 10 I, prior 100 I."""
    rng = np.random.default_rng(seed)
    Qf, _ = np.linalg.qr(rng.standard_normal((d, d)))
    return dict(A=0.99 * Qf, B=np.eye(d), P=0.05 * np.eye(d), Q=10.0 * np.eye(d),
                m0=np.zeros(d), S0=100.0 * np.eye(d))


def generate_data(model, T, batch, seed=42, chain_offset=0):
    """Generative loop of ipynb:134-148 (x0 = 0, x_t = A x_{t-1} + N(0,P), y_t = B x_t + N(0,Q))
    with NumPy ``default_rng(SeedSequence([seed, chain]))`` per chain; generated in fp64 and
    rounded to fp32 ONCE -- oracle and GPU both consume the fp32-rounded values.
    Returns (x[T, d, batch] fp64, y[T, m, batch] fp32)."""
    A, B, P, Q = model["A"], model["B"], model["P"], model["Q"]
    d = A.shape[0]; m = B.shape[0]
    LP = np.linalg.cholesky(P); LQ = np.linalg.cholesky(Q)
    x = np.zeros((T, d, batch)); y = np.zeros((T, m, batch))
    for b in range(batch):
        rng = np.random.default_rng(np.random.SeedSequence([seed, chain_offset + b]))
        ex = rng.standard_normal((T, d)) @ LP.T
        ey = rng.standard_normal((T, m)) @ LQ.T
        xp = np.zeros(d)
        for t in range(T):
            xp = A @ xp + ex[t]
            x[t, :, b] = xp
            y[t, :, b] = B @ xp + ey[t]
    return x, y.astype(np.float32)
