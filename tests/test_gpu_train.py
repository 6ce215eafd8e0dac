"""Training-step kernels (SURVEY.md section 8f NEXT-3) through the C ABI: forward, input gradient (dgrad) and weight
gradient (wgrad, csrc/conv_wgrad.cuh) of the convolutions against torch autograd of F.conv2d evaluated in fp64.
One fp16 pass with fp32 accumulation = TF32-grade (what the reference trains with): tolerance 3e-3 of the largest value."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from mcquic_b200 import autograd as A
from mcquic_b200.utils.synthetic import uniform

pytestmark = pytest.mark.gpu

TOL = 3e-3


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max()) / (float(b.abs().max()) + 1e-300)


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,bias,gscale", [
    (2, 16, 16, 128, 128, 3, 1, True, 1.0),       # the common block convolution
    (2, 32, 32, 64, 128, 3, 2, True, 1e-7),        # strided (ResidualBlockWithStride), tiny gradients: device-side scaling
    (3, 8, 8, 128, 128, 1, 1, True, 1.0),          # 1x1 (gate / GDN)
    (2, 16, 16, 32, 32, 3, 1, True, 1e3),          # Neon's quantizer nets: partly filled 64-channel chunks
    (1, 24, 40, 8, 32, 1, 1, False, 1.0),          # bias-free 1x1, 8 channels, ragged pixel tiles
    (2, 32, 32, 3, 64, 3, 1, True, 1.0),           # RGB stem of Neon (input padded to 8 channels; no input gradient needed)
    (2, 16, 16, 64, 3, 3, 1, True, 1e-5),          # Neon's last layer C -> 3
    (2, 32, 32, 128, 512, 3, 1, True, 1.0),        # the convolution in front of PixelShuffle
    (2, 12, 20, 256, 256, 3, 1, True, 1.0),        # 256 channels (a800_16 trunk): 2 x 2 channel tiles
    (1, 64, 64, 128, 128, 3, 2, True, 1.0),
    (5, 4, 4, 128, 128, 3, 1, True, 1.0),          # 4x4 maps: a pixel tile spans 8 images, ragged in n
])
def test_conv_function_matches_fp64_autograd(n, h, w, cin, cout, k, stride, bias, gscale):
    conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(uniform(tuple(conv.weight.shape), "train.w", 3) / (cin * k * k) ** 0.5)
        if bias:
            conv.bias.copy_(uniform((cout,), "train.b", 3) * 0.1)
    x = uniform((n, cin, h, w), "train.x", 4)
    g = uniform((n, cout, h // stride, w // stride), "train.g", 5) * gscale
    # fp64 reference
    xr = x.double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    br = conv.bias.detach().double().requires_grad_(True) if bias else None
    yr = F.conv2d(xr, wr, br, stride=stride, padding=k // 2)
    yr.backward(g.double())
    conv = conv.cuda()
    assert A.conv_supported(conv)
    xg = x.cuda().requires_grad_(True)
    before = A.train_engine().lib.mcq_kernel_launch_count()
    y = A.conv2d(conv, xg)
    assert tuple(y.shape) == tuple(yr.shape)
    y.backward(g.cuda())
    torch.cuda.synchronize()
    assert A.train_engine().lib.mcq_kernel_launch_count() - before >= 6      # split + conv, split + dgrad, wgrad + reduce
    assert _rel(y, yr.detach()) <= TOL
    assert _rel(xg.grad, xr.grad) <= TOL
    assert _rel(conv.weight.grad, wr.grad) <= TOL
    if bias:
        assert _rel(conv.bias.grad, br.grad) <= 1e-5
    assert A.train_engine().lib.mcq_device_error_flag() == 0


def test_unsupported_convolutions_use_the_library_path_and_still_differentiate():
    conv = nn.Conv2d(32, 32, 3, stride=2, padding=1).cuda()       # stride 2 with a 32-channel tap view
    assert not A.conv_supported(conv)
    x = uniform((2, 32, 16, 16), "train.x32", 1).cuda().requires_grad_(True)
    A.conv2d(conv, x).sum().backward()
    assert x.grad is not None and conv.weight.grad is not None
