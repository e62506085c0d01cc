"""``infer()``: the entry point of the reference (/root/reference/src/inference/inference.jl:577-733)
mirrored for the batched Gaussian hot path.

In the reference ``infer`` builds a factor graph from an ``@model`` and lets ReactiveMP/Rocket
schedule one message at a time (src/inference/batch.jl:103-482, streaming.jl:536-845).  Here the
``model`` argument is the *recognised pattern* -- what the Julia-side shim (julia/RxGaussB200.jl)
extracts from the GraphPPL graph -- and ``data`` carries ``batch`` independent series at once.
The keyword surface, the result object and the error behaviour follow the reference; every
keyword that would need machinery outside the hot path raises ``NotImplementedError`` instead of
being silently ignored (those calls must be routed to stock ReactiveMP).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .context import Context
from .distributions import GammaShapeRate, MvNormalMeanCovariance, NormalMeanVariance


# --------------------------------------------------------------------------- recognised models
@dataclass
class linear_gaussian_ssm_smoothing:
    """``@model linear_gaussian_ssm_smoothing(y, A, B, P, Q)``
    (/root/reference/benchmarks/Linear Multivariate Gaussian State Space Model Benchmark.ipynb:95-105;
    same graph as test/models/statespace/mlgssm_test.jl:8-17 with the prior on x[1]).
    ``x0 = (mean, cov)`` of the prior on x[1]."""
    A: np.ndarray
    B: np.ndarray
    P: np.ndarray
    Q: np.ndarray
    x0: tuple
    per_chain: bool = False     # model matrices carry a trailing [batch] axis (CUDA tensors)
    u: object = None            # constant offset: x[t] ~ MvNormal(A x[t-1] + u, P)  (`+` with a PointMass)
    prior_on_previous_state: bool = False   # x_prior ~ x0; x[1] ~ N(A x_prior + u, P)  (mlgssm_test.jl:8-17)


@dataclass
class linear_gaussian_ssm_filtering(linear_gaussian_ssm_smoothing):
    """One-step model + ``@autoupdates x_min_t_mean, x_min_t_cov = mean_cov(q(x_t))``
    (ipynb:107-113, 199-216): the prior is pushed through (A, P) before every datum."""


@dataclass
class hgf:
    """``@model hgf`` with ``hgfconstraints`` / ``hgfmeta`` / autoupdates of
    /root/reference/test/models/statespace/hgf_tests.jl:10-69."""
    real_k: float = 1.0
    real_w: float = 0.0
    z_variance: float = 0.2 ** 2
    y_variance: float = 0.1 ** 2
    init: tuple = (0.0, 5.0, 0.0, 5.0)      # q(zt) = N(0,5), q(xt) = N(0,5)  (hgf_tests.jl:51-54)


@dataclass
class univariate_lgssm_gamma_precision:
    """Scalar random walk observed with unknown precision tau ~ Gamma(a0, b0), q(x) q(tau)
    (rules of /root/reference/test/models/aliases/aliases_gamma_tests.jl)."""
    a: float = 1.0
    v_proc: float = 1.0
    x0: tuple = (0.0, 100.0)
    gamma_prior: tuple = (1.0, 1.0)
    init_E_tau: float = 1.0


@dataclass
class kalman_gamma_streaming:
    """The reference's ``test_model1`` (/root/reference/test/inference/inference_tests.jl:752-775): one-step random
    walk observed with unknown precision, ``constraints = MeanField()``, ``@autoupdates`` of ``mean_var(q(x_t))`` and
    ``shape / rate of q(τ)``; ``init = (m_x, v_x, shape, rate)`` of the ``@initialization`` (:766-769)."""
    transition_precision: float = 1.0
    init: tuple = (0.0, 1e3, 1.0, 1.0)


@dataclass
class latent_autoregressive:
    """``lar_model`` with ``lar_constraints`` / ``lar_init_marginals``
    (/root/reference/test/models/autoregressive/lar_tests.jl:52-122): gamma ~ Gamma, theta ~ N(0, I / w0), x0 ~ N(0, I / p0),
    x[t] ~ AR(x[t-1], theta, gamma) with ARMeta(variate, order, ARsafe()), y[t] ~ Normal(dot(c, x[t]), 1 / tau), c = e1,
    q(x, x0) q(gamma) q(theta).  The Univariate case is order = 1."""
    order: int
    tau: float
    gamma_prior: tuple = (1.0, 1.0)
    theta_prior_precision: float = 1.0
    x0_prior_precision: float = 1.0
    init_gamma: tuple = (1.0, 1.0)          # q(gamma) of the @initialization
    init_theta_precision: float = 1.0       # q(theta) = N(0, I / init_theta_precision)


@dataclass
class InferenceResult:
    """``InferenceResult`` (/root/reference/src/inference/batch.jl:18-24)."""
    posteriors: dict
    free_energy: object = None
    model: object = None
    error: object = None
    history: dict = field(default_factory=dict)


_UNSUPPORTED = ("constraints", "meta", "callbacks", "annotations", "predictvars", "events", "uselock",
                "postprocess", "trace", "benchmark", "free_energy_diagnostics")
_ctx_cache: dict = {}


def default_context(device=None) -> Context:
    dev = torch.cuda.current_device() if device is None else device
    if dev not in _ctx_cache:
        _ctx_cache[dev] = Context(dev)
    ctx = _ctx_cache[dev]
    ctx.bind_stream()
    return ctx


def infer(*, model, iterations=None, free_energy=False, returnvars=None, options=None,
          initialization=None, autoupdates=None, keephistory=None, historyvars=None,
          catch_exception=False, showprogress=False, session=None, warn=True, allow_node_contraction=False,
          context: Context | None = None, cov_shared_out=False, data=None, datastream=None, autostart=True,
          batch=None, **kwargs):
    """Batched ``infer``.  ``data = {"y": tensor[T, m, batch]}`` (CUDA fp32, or CPU for the
    host-staged path).  Returns ``posteriors["x"]`` as a batched ``MvNormalMeanCovariance``.

    With ``datastream=`` (an iterable of time-chunks, or ``None`` + ``autoupdates`` for a push-driven
    engine) the call returns an ``RxInferenceEngine`` (streaming.py), as the reference does when
    ``autoupdates`` is given (/root/reference/src/inference/inference.jl:577-733 dispatch)."""
    for k in kwargs:
        if k in _UNSUPPORTED:
            raise NotImplementedError(
                f"infer(..., {k}=...) needs per-message machinery outside the batched hot path; "
                "run this call through stock ReactiveMP")
        raise TypeError(f"infer() got an unexpected keyword argument '{k}'")
    if options:
        bad = set(options) - {"limit_stack_depth", "warn"}   # limit_stack_depth is moot: the schedule is a fused sweep
        if bad:
            raise NotImplementedError(f"options {sorted(bad)} are outside the batched hot path")
    if data is not None and datastream is not None:
        raise ValueError("`data` and `datastream` are mutually exclusive")    # reference: inference.jl argument check
    if data is None:
        if datastream is None and autoupdates is None:
            raise ValueError("either `data` or `datastream` (or `autoupdates` for a push-driven engine) is required")
        if batch is None:
            raise ValueError("streaming inference needs `batch` (number of lock-step datastreams)")
        from .streaming import RxInferenceEngine
        return RxInferenceEngine(context or default_context(), model, batch=batch, iterations=iterations,
                                 keephistory=keephistory, historyvars=historyvars, free_energy=free_energy,
                                 datastream=datastream, autostart=autostart, cov_shared_out=cov_shared_out)
    if "y" not in data:
        raise KeyError("data must contain the observations under key 'y'")   # reference: missing data key error
    ctx = context or default_context()
    y = data["y"]
    mask = data.get("ymask")
    try:
        if isinstance(model, linear_gaussian_ssm_filtering):
            r = ctx.lgssm(y, model.A, model.B, model.P, model.Q, model.x0[0], model.x0[1], u=model.u, smooth=False, mask=mask,
                          want_evidence=free_energy, per_chain_model=model.per_chain, transition_first=True,
                          cov_shared_out=cov_shared_out)
            q = MvNormalMeanCovariance(r["mean"], r["cov"])
            return InferenceResult(posteriors={}, history={"x_t": q}, free_energy=r["neg_log_evidence"], model=model)
        if isinstance(model, linear_gaussian_ssm_smoothing):
            if iterations not in (None, 1):
                raise NotImplementedError("iterations > 1 on a tree-structured BP model is a no-op in the reference; "
                                          "KeepEach() results are outside the hot path")
            r = ctx.lgssm(y, model.A, model.B, model.P, model.Q, model.x0[0], model.x0[1], u=model.u, smooth=True, mask=mask,
                          want_evidence=free_energy, per_chain_model=model.per_chain, cov_shared_out=cov_shared_out,
                          transition_first=model.prior_on_previous_state)
            return InferenceResult(posteriors={"x": MvNormalMeanCovariance(r["mean"], r["cov"])},
                                   free_energy=r["neg_log_evidence"], model=model)
        if isinstance(model, hgf):
            out = ctx.hgf_filter(y, iters=iterations or 1, kappa=model.real_k, omega=model.real_w,
                                 z_variance=model.z_variance, y_variance=model.y_variance, init=model.init,
                                 want_free_energy=bool(free_energy))
            fe = None
            if free_energy:      # free_energy_history of the streaming engine: average over the data, per iteration
                out, fe_all = out
                fe = fe_all.mean(dim=0)
            return InferenceResult(posteriors={}, model=model, free_energy=fe,
                                   history={"xt": NormalMeanVariance(out[:, 0], out[:, 1]),
                                            "zt": NormalMeanVariance(out[:, 2], out[:, 3])})
        if isinstance(model, kalman_gamma_streaming):
            out, fe = ctx.stream_vmp_gamma(y, iters=iterations or 1, w=model.transition_precision, init=model.init,
                                           want_free_energy=bool(free_energy))
            return InferenceResult(posteriors={}, model=model, free_energy=None if fe is None else fe.mean(dim=0),
                                   history={"x_t": NormalMeanVariance(out[:, 0], out[:, 1]),
                                            "τ": GammaShapeRate(out[:, 2], out[:, 3])})
        if isinstance(model, univariate_lgssm_gamma_precision):
            r = ctx.lgssm_vmp_gamma(y, iterations=iterations or 1, a=model.a, v_proc=model.v_proc, prior=model.x0,
                                    gamma_prior=model.gamma_prior, init_E_tau=model.init_E_tau,
                                    want_free_energy=bool(free_energy))
            return InferenceResult(posteriors={"x": NormalMeanVariance(r["mean"], r["var"]),
                                               "τ": GammaShapeRate(r["shape"], r["rate"])}, model=model,
                                   free_energy=r["free_energy"])
        if isinstance(model, latent_autoregressive):
            yy = y[:, 0] if y.dim() == 3 else y
            r = ctx.lar_vmp(yy.contiguous(), model.order, model.tau, iterations=iterations or 1, gamma_prior=model.gamma_prior,
                            theta_prior_precision=model.theta_prior_precision, x0_prior_precision=model.x0_prior_precision,
                            init_gamma=model.init_gamma, init_theta_precision=model.init_theta_precision,
                            want_free_energy=bool(free_energy))
            # returnvars of the reference's call: x = KeepLast(), gamma / theta = KeepEach() (leading iteration axis)
            return InferenceResult(posteriors={"x": MvNormalMeanCovariance(r["x_mean"], r["x_cov"]),
                                               "γ": GammaShapeRate(r["gamma_shape"], r["gamma_rate"]),
                                               "θ": MvNormalMeanCovariance(r["theta_mean"], r["theta_cov"])},
                                   model=model, free_energy=r["free_energy"])
        raise NotImplementedError(f"model pattern {type(model).__name__} is not on the batched hot path")
    except Exception as e:           # reference: catch_exception=true returns a partial result with .error
        if catch_exception:
            return InferenceResult(posteriors={}, model=model, error=e)
        raise
