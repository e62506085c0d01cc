"""fp64 oracle for the Hierarchical Gaussian Filter streaming path (GCV node, GH-31).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED at posterior level: the reference pins this model only through
StableRNG-generated data (/root/reference/test/models/statespace/hgf_tests.jl:94-133), which
cannot be regenerated without Julia.  What IS anchored: the model/constraints/meta/autoupdates
(hgf_tests.jl:10-69) and the in-repo restatement of A, B, ksi, psi and the node energy
(/root/reference/test/inference/inference_tests.jl:587-607).  Update order inside one VMP
iteration is the dependency-respecting one of SURVEY.md appendix A.4.
"""
from __future__ import annotations

import numpy as np

from . import rules as R


def hgf_filter(y, iters=20, kappa=1.0, omega=0.0, z_variance=0.2 ** 2, y_variance=0.1 ** 2,
               init=(0.0, 5.0, 0.0, 5.0), return_free_energy=False):
    """y[T, batch] -> out[T, 4, batch] = (m_x, v_x, m_z, v_z) per step
    (= ``history[:xt]``, ``history[:zt]`` of hgf_tests.jl:106-107).

    Per datum (streaming engine loop /root/reference/src/inference/streaming.jl:349-407):
    priors (z-, x-) are the previous step's q(zt), q(xt) (``@autoupdates`` hgf_tests.jl:46-49),
    q(z) carried across iterations starts from the previous posterior (init N(0,5),
    hgf_tests.jl:51-54)."""
    y = np.asarray(y, dtype=np.float64)
    T, batch = y.shape
    nw = R.gauss_hermite(31)
    mzp = np.full(batch, init[0]); vzp = np.full(batch, init[1])
    mxp = np.full(batch, init[2]); vxp = np.full(batch, init[3])
    qz = (mzp.copy(), vzp.copy())
    out = np.zeros((T, 4, batch))
    fe = np.zeros((T, iters, batch)) if return_free_energy else None
    for t in range(T):
        m_y = (y[t], np.full(batch, y_variance))           # NormalMeanVariance(:mu) from data
        m_x = (mxp, vxp)                                   # prior on xt_min
        z_prior = R.normal_meanvar_out((mzp, vzp), z_variance)   # zt ~ N(zt_min, z_variance)
        for it in range(iters):
            m, V = R.gcv_marginal_yx(m_y, m_x, qz, kappa, omega)
            elq = R.gcv_z_elq(m, V, kappa, omega)
            qz = R.prod_normal_elq(z_prior, elq, nw)
            if return_free_energy:
                fe[t, it] = _hgf_step_energy(m, V, qz, kappa, omega)
        # history holds q(xt) from the last firing of the joint marginal (appendix A.4 order:
        # joint -> psi -> q(z)), i.e. computed with the q(z) of the previous iteration
        out[t, 0], out[t, 1] = m[..., 0], V[..., 0, 0]
        out[t, 2], out[t, 3] = qz
        mxp, vxp = m[..., 0], V[..., 0, 0]
        mzp, vzp = qz
    if return_free_energy:
        return out, fe
    return out


def _hgf_step_energy(m, V, qz, kappa, omega):
    """GCV node average energy U = 1/2 [log 2pi + (kappa m_z + omega) + psi A B]
    (/root/reference/test/inference/inference_tests.jl:595-606)."""
    mz, vz = qz
    psi = (m[..., 0] - m[..., 1]) ** 2 + V[..., 0, 0] + V[..., 1, 1] - 2 * V[..., 0, 1]
    A = np.exp(-omega)
    B = np.exp(-kappa * mz + 0.5 * kappa ** 2 * vz)
    return 0.5 * (np.log(2 * np.pi) + (kappa * mz + omega) + psi * A * B)


def generate_data(T, batch, kappa=1.0, omega=0.0, z_variance=0.2 ** 2, y_variance=0.1 ** 2, seed=42):
    """Generative loop of hgf_tests.jl:72-92, vectorised over chains; y rounded to fp32 once."""
    rng = np.random.default_rng(np.random.SeedSequence([seed, 7]))
    z = np.zeros((T, batch)); x = np.zeros((T, batch)); y = np.zeros((T, batch))
    zp = np.zeros(batch); xp = np.zeros(batch)
    for t in range(T):
        zp = zp + np.sqrt(z_variance) * rng.standard_normal(batch)
        v = np.exp(kappa * zp + omega)
        xp = xp + np.sqrt(v) * rng.standard_normal(batch)
        z[t], x[t] = zp, xp
        y[t] = xp + np.sqrt(y_variance) * rng.standard_normal(batch)
    return z, x, y.astype(np.float32)
