"""GDN / inverse GDN parameter containers (mcquic/nn/gdn.py:28-91, mcquic/nn/base.py:31-84).

Same parameters and persistent buffers as the reference, so reference checkpoints load unchanged:
  beta [C], gamma [C, C], {beta,gamma}_reparam.eps [1], {beta,gamma}_reparam.lowerBound.bound [1].
The non-negative reparametrisation is folded once per weight version (`effective()`), instead of on every
forward as the reference does.
"""
import torch
from torch import nn

__all__ = ["GenDivNorm", "InvGenDivNorm", "NonNegativeParametrizer", "LowerBound"]

_EPS = 1e-6  # mcquic/consts.py:25


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return torch.max(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum: float = 0.0, eps: float = _EPS):
        super().__init__()
        self.register_buffer("eps", torch.Tensor([float(eps) ** 2]))
        self.lowerBound = LowerBound((float(minimum) + float(eps) ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.eps, self.eps))

    def forward(self, x):
        return self.lowerBound(x) ** 2 - self.eps


class GenDivNorm(nn.Module):
    inverse = False

    def __init__(self, inChannels: int, groups: int = 1, biasBound: float = 1e-4, weightInit: float = 0.1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("mcquic_b200: grouped GDN is not on the accelerated path")
        self.beta_reparam = NonNegativeParametrizer(minimum=float(biasBound))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(inChannels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(weightInit) * torch.eye(inChannels)))

    @torch.no_grad()
    def effective(self):
        """(beta_eff [C], gamma_eff [C, C]) as used by y = x * rsqrt(beta + gamma (*) x^2)."""
        return self.beta_reparam(self.beta), self.gamma_reparam(self.gamma)


class InvGenDivNorm(GenDivNorm):
    inverse = True
