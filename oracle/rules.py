"""fp64 NumPy restatement of the reference's Gaussian-family message-update rules.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function is the body of one ``@rule`` / ``prod`` of the reference stack, batched
over arbitrary leading axes (``...``); vectors are ``[..., d]``, matrices ``[..., d, d]``.
Parametrisations: (mu, Sigma) = mean / covariance (``MvNormalMeanCovariance``),
(xi, W) = weighted mean / precision (``MvNormalWeightedMeanPrecision``), W = inv(Sigma),
xi = W mu.  Constants and data arrive as PointMass because they are factorised out by
default (/root/reference/src/model/model.jl:198,222).

The rule bodies themselves live in ReactiveMP ~6.0.0 / ExponentialFamily 2.1.0 which are not
vendored under /root/reference (Project.toml:49,63); each function cites the reference
call site that exercises it and the upstream file it restates (marked "upstream").
"""
from __future__ import annotations

import numpy as np

TINY = 1e-12   # TinyHugeNumbers.tiny (upstream), used by vague(GammaShapeRate)
HUGE = 1e12    # TinyHugeNumbers.huge (upstream), used by vague(NormalMeanVariance)


# --------------------------------------------------------------------------- linear algebra
def cholinv(S):
    """SPD inverse through a Cholesky factorisation.

    Restates ``FastCholesky.cholinv`` (FastCholesky 1.3.0, re-exported at
    /root/reference/src/RxInfer.jl:6); used by every (mu,Sigma) <-> (xi,W) conversion.
    """
    S = np.asarray(S, dtype=np.float64)
    L = np.linalg.cholesky(S)
    eye = np.broadcast_to(np.eye(S.shape[-1]), S.shape)
    Linv = np.linalg.solve(L, eye)
    return np.swapaxes(Linv, -1, -2) @ Linv


def sym(M):
    return 0.5 * (M + np.swapaxes(M, -1, -2))


def mv(M, v):
    return np.einsum("...ij,...j->...i", M, v)


def meancov_to_wmp(mu, Sigma):
    """``weightedmean_precision(::MvNormalMeanCovariance)`` (upstream ExponentialFamily
    normal_family); one cholinv."""
    W = cholinv(Sigma)
    return mv(W, mu), W


def wmp_to_meancov(xi, W):
    """``mean_cov(::MvNormalWeightedMeanPrecision)``; one cholinv."""
    Sigma = cholinv(W)
    return mv(Sigma, xi), Sigma


# --------------------------------------------------------------------------- SURVEY 8a row 2
def mvnormal_meancov_out(m_mu, q_Sigma):
    """@rule MvNormalMeanCovariance(:out, Marginalisation)(m_mu::MvNormal, q_Sigma::PointMass).

    Alias ``MvNormal(mu, Sigma)`` -> node at /root/reference/src/model/graphppl.jl:372-376;
    model line benchmarks/...Benchmark.ipynb:102.  (mu_mu, Sigma_mu + Sigma).
    With a PointMass mean (the prior, ipynb:98) Sigma_mu = 0.
    """
    mu, S = m_mu
    return mu, S + q_Sigma


# --------------------------------------------------------------------------- SURVEY 8a row 3
def mvnormal_meancov_mean(m_out, q_Sigma):
    """@rule MvNormalMeanCovariance(:mu, Marginalisation)(m_out::MvNormal, q_Sigma::PointMass)
    -- backward through the process noise: (mu_out, Sigma_out + Sigma)."""
    mu, S = m_out
    return mu, S + q_Sigma


def mvnormal_meancov_mean_from_data(y, q_Sigma):
    """Same rule with q_out::PointMass (an observation pushed by
    /root/reference/src/inference/batch.jl:405-407): (y, Sigma)."""
    y = np.asarray(y, dtype=np.float64)
    return y, np.broadcast_to(q_Sigma, y.shape[:-1] + q_Sigma.shape[-2:]).copy()


# --------------------------------------------------------------------------- SURVEY 8a row 1
def multiplication_out(A, m_in):
    """@rule typeof(*)(:out, Marginalisation)(m_A::PointMass{Matrix}, m_in::MvNormal).

    ``A * x[t-1]`` at benchmarks/...ipynb:102, test/models/statespace/mlgssm_test.jl:13;
    (A mu, A Sigma A').
    """
    mu, S = m_in
    At = np.swapaxes(A, -1, -2)
    return mv(A, mu), A @ S @ At


# --------------------------------------------------------------------------- SURVEY 8a row 4
def multiplication_in(m_out_wmp, A):
    """@rule typeof(*)(:in, Marginalisation)(m_out::MvNormal, m_A::PointMass{Matrix}, meta).

    Returns the (xi, W) message (A' xi_out, A' W_out A); the operand is converted with
    ``weightedmean_precision`` first (one cholinv when it arrives as (mu, Sigma)).
    """
    xi, W = m_out_wmp
    At = np.swapaxes(A, -1, -2)
    return mv(At, xi), At @ W @ A


# --------------------------------------------------------------------------- SURVEY 8a row 5
def addition_out(m_in1, m_in2):
    """@rule typeof(+)(:out, Marginalisation)(m_in1, m_in2): (mu1+mu2, S1+S2).
    Call site test/models/statespace/ulgssm_tests.jl:12."""
    return m_in1[0] + m_in2[0], m_in1[1] + m_in2[1]


def addition_in1(m_out, m_in2):
    """@rule typeof(+)(:in1)(m_out, m_in2): (mu_out - mu2, S_out + S2)."""
    return m_out[0] - m_in2[0], m_out[1] + m_in2[1]


def addition_in2(m_out, m_in1):
    """@rule typeof(+)(:in2)(m_out, m_in1): (mu_out - mu1, S_out + S1)."""
    return m_out[0] - m_in1[0], m_out[1] + m_in1[1]


# --------------------------------------------------------------------------- SURVEY 8a rows 6-7
def prod_gaussian_wmp(l, r):
    """``BayesBase.prod(::GenericProd, ::MvNormal, ::MvNormal)``: (xi1+xi2, W1+W2);
    folded left-to-right at /root/reference/src/model/plugins/reactivemp_inference.jl:365-374."""
    return l[0] + r[0], l[1] + r[1]


def marginal_from_messages(msgs_wmp):
    """Product of all inbound (xi, W) messages, then ``mean_cov`` (one more cholinv)."""
    xi = sum(m[0] for m in msgs_wmp)
    W = sum(m[1] for m in msgs_wmp)
    return wmp_to_meancov(xi, W)


# --------------------------------------------------------------------------- SURVEY 8a row 8
def normal_meanvar_out(m_mu, v):
    """@rule NormalMeanVariance(:out)(m_mu::Normal, q_v::PointMass): (m, v_mu + v).
    Alias at /root/reference/src/model/graphppl.jl:340-370."""
    return m_mu[0], m_mu[1] + v


def normal_meanprec_out_q_tau(m_mu, E_tau):
    """@rule NormalMeanPrecision(:out)(m_mu::UnivariateNormalDistributionsFamily, q_tau): the BP message on the
    mean edge keeps its variance: (m_mu, v_mu + 1/E[tau])."""
    return m_mu[0], m_mu[1] + 1.0 / E_tau


def normal_meanprec_out_q_mu_q_tau(q_mu, E_tau):
    """@rule NormalMeanPrecision(:out)(q_mu::Any, q_tau::Any) (mean-field):
    NormalMeanPrecision(mean(q_mu), mean(q_tau)), i.e. variance 1/E[tau] only -- var(q_mu) does not enter."""
    return q_mu[0], 0.0 * np.asarray(q_mu[1]) + 1.0 / E_tau


def prod_normal_mv(l, r):
    """Univariate Gaussian product in (mean, var) I/O (precision-weighted combine)."""
    w = 1.0 / l[1] + 1.0 / r[1]
    xi = l[0] / l[1] + r[0] / r[1]
    return xi / w, 1.0 / w


# --------------------------------------------------------------------------- SURVEY 8a row 9
def normal_meanprec_tau(q_out, q_mu):
    """@rule NormalMeanPrecision(:tau)(q_out, q_mu) (mean-field):
    GammaShapeRate(3/2, 1/2 [(m_out - m_mu)^2 + v_out + v_mu]).
    Exercised by test/models/aliases/aliases_gamma_tests.jl:13-18."""
    m1, v1 = q_out
    m2, v2 = q_mu
    return 1.5 + 0.0 * np.asarray(m1), 0.5 * ((m1 - m2) ** 2 + v1 + v2)


def normal_meanprec_tau_structured(m, V):
    """@rule NormalMeanPrecision(:tau)(q_out_mu) with a joint (out, mu) marginal:
    GammaShapeRate(3/2, 1/2 [V11 + V22 - V12 - V21 + (m1 - m2)^2])."""
    return 1.5 + 0.0 * m[..., 0], 0.5 * (
        V[..., 0, 0] + V[..., 1, 1] - V[..., 0, 1] - V[..., 1, 0] + (m[..., 0] - m[..., 1]) ** 2
    )


def mvnormal_meanprec_lambda(q_out, q_mu):
    """@rule MvNormalMeanPrecision(:Lambda)(q_out, q_mu) (mean-field): Wishart(d + 2, inv(V_out + V_mu + (m_out - m_mu)
    (m_out - m_mu)')) -- returned in the WishartFast parametrisation (df, INVERSE scale).  Exercised by
    /root/reference/test/models/iid/mv_iid_precision_tests.jl:10-16 (upstream rules/mv_normal_mean_precision/precision.jl)."""
    (mo, Vo), (mm, Vm) = q_out, q_mu
    dlt = mo - mm
    d = mo.shape[-1]
    return d + 2.0 + 0.0 * mo[..., 0], Vo + Vm + dlt[..., :, None] * dlt[..., None, :]


def prod_wishart(l, r):
    """prod(Wishart(nu1, S1), Wishart(nu2, S2)) = Wishart(nu1 + nu2 - d - 1, inv(inv(S1) + inv(S2))); (df, inverse scale) I/O."""
    d = l[1].shape[-1]
    return l[0] + r[0] - d - 1.0, l[1] + r[1]


def wishart_mean(w):
    """mean(Wishart(df, S)) = df S, with S = inv(inverse scale)."""
    return w[0][..., None, None] * np.linalg.inv(w[1])


def prod_gamma(l, r):
    """``prod(GammaShapeRate, GammaShapeRate)`` = GammaShapeRate(a1 + a2 - 1, b1 + b2)."""
    return l[0] + r[0] - 1.0, l[1] + r[1]


def gamma_mean(g):
    return g[0] / g[1]


def gamma_mean_log(g):
    from scipy.special import digamma
    return digamma(g[0]) - np.log(g[1])


def gamma_entropy(g):
    from scipy.special import digamma, gammaln
    a, b = g
    return a - np.log(b) + gammaln(a) + (1.0 - a) * digamma(a)


def normal_entropy(v):
    """entropy(NormalMeanVariance(m, v)); golden 1.4189385332046727 for v = 1 at
    /root/reference/test/score/diagnostics_tests.jl:24."""
    return 0.5 * (1.0 + np.log(2.0 * np.pi * v))


def mvnormal_entropy(S):
    d = S.shape[-1]
    return 0.5 * (d * (1.0 + np.log(2.0 * np.pi)) + np.linalg.slogdet(S)[1])


# --------------------------------------------------------------------------- SURVEY 8a row 10
def gauss_hermite(n=31):
    """Nodes / weights of ``GaussHermiteCubature(31)``
    (/root/reference/test/models/statespace/hgf_tests.jl:39); physicists' convention."""
    from scipy.special import roots_hermite
    t, w = roots_hermite(n)
    return t, w


def gcv_gamma(q_z, kappa, omega):
    """A*B of the GCV node with PointMass kappa/omega:
    A = exp(-omega), B = exp(-kappa m_z + kappa^2 v_z / 2); restated in-repo at
    /root/reference/test/inference/inference_tests.jl:595-606 (ksi, A, B)."""
    mz, vz = q_z
    ksi = kappa ** 2 * vz
    return np.exp(-omega) * np.exp(-kappa * mz + 0.5 * ksi)


def gcv_y(m_x, q_z, kappa, omega):
    """@rule GCV(:y)(m_x, q_z, q_kappa, q_omega): N(m_x, v_x + 1/(A B))."""
    return m_x[0], m_x[1] + 1.0 / gcv_gamma(q_z, kappa, omega)


def gcv_x(m_y, q_z, kappa, omega):
    """@rule GCV(:x) -- symmetric to :y."""
    return m_y[0], m_y[1] + 1.0 / gcv_gamma(q_z, kappa, omega)


def gcv_marginal_yx(m_y, m_x, q_z, kappa, omega):
    """@marginalrule GCV(:y_x): joint MvNormalWeightedMeanPrecision with
    xi = [xi_y, xi_x], W = [[w_y + g, -g], [-g, w_x + g]], g = A B.  Returns (m[2], V[2,2])."""
    g = gcv_gamma(q_z, kappa, omega)
    wy, wx = 1.0 / m_y[1], 1.0 / m_x[1]
    xiy, xix = m_y[0] * wy, m_x[0] * wx
    a, b, c = wy + g, -g, wx + g
    det = a * c - b * b
    V11, V12, V22 = c / det, -b / det, a / det
    m1 = V11 * xiy + V12 * xix
    m2 = V12 * xiy + V22 * xix
    m = np.stack([m1, m2], axis=-1)
    V = np.stack([np.stack([V11, V12], -1), np.stack([V12, V22], -1)], -2)
    return m, V


def gcv_z_elq(m, V, kappa, omega):
    """@rule GCV(:z)(q_y_x, q_kappa, q_omega) -> ExponentialLinearQuadratic(a, b, c, d) with
    a = kappa, b = psi * A, c = -kappa, d = v_kappa = 0;
    psi = (m_y - m_x)^2 + V_yy + V_xx - 2 V_yx."""
    psi = (m[..., 0] - m[..., 1]) ** 2 + V[..., 0, 0] + V[..., 1, 1] - 2.0 * V[..., 0, 1]
    return kappa, psi * np.exp(-omega), -kappa, 0.0


def prod_normal_elq(n, elq, nodes_weights=None):
    """prod(Normal, ExponentialLinearQuadratic) by Gauss-Hermite moment matching around the
    Gaussian factor (upstream approximations/gausshermite.jl).  ELQ density is
    exp(-(a z + b exp(c z + d z^2 / 2)) / 2)."""
    t, w = nodes_weights if nodes_weights is not None else gauss_hermite(31)
    m0, v0 = (np.asarray(n[0], dtype=np.float64), np.asarray(n[1], dtype=np.float64))
    a, b, c, d = elq
    z = m0[..., None] + np.sqrt(2.0 * v0)[..., None] * t
    bb = np.asarray(b, dtype=np.float64)[..., None]
    g = np.exp(-0.5 * (a * z + bb * np.exp(c * z + 0.5 * d * z * z)))
    wg = w * g
    Z = wg.sum(-1)
    mz = (wg * z).sum(-1) / Z
    vz = (wg * (z - mz[..., None]) ** 2).sum(-1) / Z
    return mz, vz


def elq_mean_var(elq, nodes_weights=None):
    """``mean_var(::ExponentialLinearQuadratic)`` as the reference evaluates it: moments of the (integrable) ELQ density
    by ``GaussHermiteCubature(31)`` against a standard normal with the pdf re-weighted by exp(z^2 / 2)
    (``approximate_meancov(approximation, adjusted_pdf, NormalMeanVariance(0, 1))``, upstream
    distributions/exp_linear_quadratic.jl).  This is how a Normal node treats an inbound ELQ message in its
    (out, mu) marginal rule; reproducing the reference's HGF free energy 1.009879989585
    (/root/reference/test/models/statespace/hgf_tests.jl:118) to 1e-5 pins this reading (exact moments of the ELQ give
    1.00704, the q(zt)/forward-message Gaussian 1.05178)."""
    t, w = nodes_weights if nodes_weights is not None else gauss_hermite(31)
    a, b, c, d = elq
    z = np.sqrt(2.0) * t
    bb = np.asarray(b, dtype=np.float64)[..., None]
    g = np.exp(-0.5 * (a * z + bb * np.exp(c * z + 0.5 * d * z * z)) + 0.5 * z * z)
    wg = w * g
    Z = wg.sum(-1)
    m = (wg * z).sum(-1) / Z
    v = (wg * (z - m[..., None]) ** 2).sum(-1) / Z
    return m, v
