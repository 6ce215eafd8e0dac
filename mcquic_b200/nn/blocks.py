"""Residual / attention blocks with the reference's constructor signatures and state_dict layout
(mcquic/nn/blocks.py:62-78 _residulBlock, :82-122 ResidualBlockWithStride, :125-159 ResidualBlockShuffle,
:163-200 ResidualBlock, :246-288 AttentionBlock).

`_branch` is the same 4-slot Sequential as upstream (activation, conv, activation/norm, conv) so that the keys
`_branch.1.*`, `_branch.2.*`, `_branch.3.*`, `_skip.*` line up.  `forward` runs the block on the CUDA
engine (NCHW fp32 in / out) -- there is no PyTorch arithmetic fallback.
"""
from typing import Optional

import torch
from torch import nn

from .convs import conv1x1, conv3x3, pixelShuffle3x3
from .gdn import GenDivNorm, InvGenDivNorm

__all__ = ["ResidualBlock", "ResidualBlockWithStride", "ResidualBlockShuffle", "AttentionBlock"]


class _EngineModule(nn.Module):
    """Stand-alone call of a block: convert at the boundary, run on the engine."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..engine import default_engine
        return default_engine().run_module_nchw(self, x)


class _ResidualBase(_EngineModule):
    def __init__(self, act1: nn.Module, conv1: nn.Module, act2: nn.Module, conv2: nn.Module, skip: Optional[nn.Module]):
        super().__init__()
        self._branch = nn.Sequential(act1, conv1, act2, conv2)
        self._skip = skip


class ResidualBlock(_ResidualBase):
    """y = skip(x) + conv3(act2(conv3(SiLU(x)))),  act2 = SiLU, or nn.GroupNorm(groups, outChannels) when denseNorm
    (then NO activation before the second conv, blocks.py:198); skip = conv1x1 iff the channel count changes
    (blocks.py:189-192), else the identity.  `groups` is the GroupNorm group count, not a grouped convolution."""

    def __init__(self, inChannels: int, outChannels: int, groups: int = 1, denseNorm: bool = False):
        skip = conv1x1(inChannels, outChannels) if inChannels != outChannels else None
        super().__init__(nn.SiLU(), conv3x3(inChannels, outChannels),
                         nn.GroupNorm(groups, outChannels) if denseNorm else nn.SiLU(),
                         conv3x3(outChannels, outChannels), skip)


class ResidualBlockWithStride(_ResidualBase):
    """y = conv3s2_skip(x) + conv3(GDN(conv3s2(SiLU(x))))      (H -> H/2)"""

    def __init__(self, inChannels: int, outChannels: int, stride: int = 2, groups: int = 1, denseNorm: bool = False):
        if stride != 2:
            raise NotImplementedError("mcquic_b200: ResidualBlockWithStride is accelerated for stride 2")
        super().__init__(nn.SiLU(), conv3x3(inChannels, outChannels, stride=2), GenDivNorm(outChannels),
                         conv3x3(outChannels, outChannels), conv3x3(inChannels, outChannels, stride=2))


class ResidualBlockShuffle(_ResidualBase):
    """y = PS2(conv3_skip(x)) + conv3(IGDN(PS2(conv3(SiLU(x)))))      (H -> 2H)"""

    def __init__(self, inChannels: int, outChannels: int, upsample: int = 2, groups: int = 1, denseNorm: bool = False):
        super().__init__(nn.SiLU(), pixelShuffle3x3(inChannels, outChannels, upsample), InvGenDivNorm(outChannels),
                         conv3x3(outChannels, outChannels), pixelShuffle3x3(inChannels, outChannels, upsample))


class AttentionBlock(_EngineModule):
    """y = x + main(x) * sigmoid(side(x)); main = 3 ResidualBlocks, side = 3 ResidualBlocks + conv1x1"""

    def __init__(self, channel: int, groups: int = 1, denseNorm: bool = False):
        super().__init__()
        self._mainBranch = nn.Sequential(*[ResidualBlock(channel, channel, groups, denseNorm) for _ in range(3)])
        self._sideBranch = nn.Sequential(*[ResidualBlock(channel, channel, groups, denseNorm) for _ in range(3)],
                                         conv1x1(channel, channel))
