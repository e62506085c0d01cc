"""Batched message / marginal containers named after the reference's distribution types
(ExponentialFamily.jl, aliases at /root/reference/src/model/graphppl.jl:340-423).

Every container holds structure-of-arrays CUDA tensors with the batch axis innermost, the layout
of the C ABI.  ``mean`` / ``cov`` / ``var`` / ``mean_cov`` mirror the accessors user code calls
on ``posteriors[:x]`` (e.g. /root/reference/test/models/statespace/mlgssm_test.jl:121-126).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class PointMass:
    """Constant / datum (factorised out by default, /root/reference/src/model/model.jl:198,222)."""
    value: object


@dataclass
class MvNormalMeanCovariance:
    mu: torch.Tensor       # [..., d, n]
    Sigma: torch.Tensor    # [..., d, d, n]  (or [..., d, d] when shared across the batch)

    def mean(self):
        return self.mu

    def cov(self):
        return self.Sigma

    def var(self):
        return torch.diagonal(self.Sigma, dim1=-3, dim2=-2) if self.Sigma.dim() == self.mu.dim() + 1 \
            else torch.diagonal(self.Sigma, dim1=-2, dim2=-1)

    def mean_cov(self):
        return self.mu, self.Sigma


@dataclass
class MvNormalWeightedMeanPrecision:
    xi: torch.Tensor
    W: torch.Tensor

    def weightedmean_precision(self):
        return self.xi, self.W


@dataclass
class NormalMeanVariance:
    m: torch.Tensor
    v: torch.Tensor

    def mean(self):
        return self.m

    def var(self):
        return self.v

    def mean_var(self):
        return self.m, self.v


@dataclass
class GammaShapeRate:
    a: torch.Tensor
    b: torch.Tensor

    def mean(self):
        return self.a / self.b

    def shape(self):
        return self.a

    def rate(self):
        return self.b


@dataclass
class WishartFast:
    """Wishart in the (df, INVERSE scale) parametrisation ReactiveMP uses for messages (``WishartFast``):
    ``df[n]``, ``invS[d, d, n]``; products are additions."""
    df: torch.Tensor
    invS: torch.Tensor


def vague(kind, like: torch.Tensor):
    """``vague(NormalMeanVariance)`` = N(0, 1e12); ``vague(GammaShapeRate)`` = Gamma(1, 1e-12)
    (TinyHugeNumbers, upstream)."""
    if kind is NormalMeanVariance:
        return NormalMeanVariance(torch.zeros_like(like), torch.full_like(like, 1e12))
    if kind is GammaShapeRate:
        return GammaShapeRate(torch.ones_like(like), torch.full_like(like, 1e-12))
    raise TypeError(kind)
