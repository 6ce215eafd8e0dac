"""Convolution factories with the reference's names, signatures and state_dict layout
(mcquic/nn/convs.py:77-100 conv3x3, :221-255 pixelShuffle3x3, :257-276 conv1x1).

The modules are parameter containers: the arithmetic runs in libmcquic_b200.so through
`mcquic_b200.engine.Engine`, which recognises them by type.
"""
from torch import nn

__all__ = ["conv3x3", "conv1x1", "pixelShuffle3x3"]


def conv3x3(inChannels: int, outChannels: int, stride: int = 1, bias: bool = True, groups: int = 1) -> nn.Conv2d:
    if groups != 1 or not bias:
        raise NotImplementedError("mcquic_b200: only groups=1, bias=True convolutions are on the accelerated path")
    return nn.Conv2d(inChannels, outChannels, kernel_size=3, stride=stride, padding=1)


def conv1x1(inChannels: int, outChannels: int, stride: int = 1, bias: bool = True, groups: int = 1) -> nn.Conv2d:
    if groups != 1 or stride != 1:
        raise NotImplementedError("mcquic_b200: only groups=1, stride=1 1x1 convolutions are accelerated")
    return nn.Conv2d(inChannels, outChannels, kernel_size=1, bias=bias)


def pixelShuffle3x3(inChannels: int, outChannels: int, r: float = 1, groups: int = 1) -> nn.Sequential:
    """conv3x3(C -> r*r*Cout) followed by PixelShuffle(r); state_dict keys `0.weight`, `0.bias`."""
    if groups != 1 or int(r) != 2:
        raise NotImplementedError("mcquic_b200: only the 2x up-sampling pixelShuffle3x3 is on the accelerated path")
    return nn.Sequential(nn.Conv2d(inChannels, outChannels * 4, kernel_size=3, padding=1), nn.PixelShuffle(2))
