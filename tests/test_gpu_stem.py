"""The analysis transform's first layer (conv3x3 stride 2, RGB -> C, compressor.py:124) through the C ABI: the tcgen05 kernel
(csrc/stem_tc.cuh, fp32-grade 3-pass split), the FFMA kernel it replaced, AlignedPadding's reflect pad folded into both
(transforms.py:86-99), and uint8 input with the reference's input transform (demo.py:110-118) folded in -- against an
fp64 evaluation of the same convolution."""
import pytest
import torch
import torch.nn.functional as F

from mcquic_b200 import _lib
from mcquic_b200.engine import Engine
from mcquic_b200.modules.compressor import aligned_pad_amounts
from mcquic_b200.nn import conv3x3
from mcquic_b200.utils.synthetic import uniform

pytestmark = pytest.mark.gpu


def _planes_value(pl):
    return pl[0].double() + (pl[1].double() / 2048.0 if pl[1] is not None else 0.0)


@pytest.mark.parametrize("n,h,w,cout", [(2, 256, 256, 128), (1, 150, 333, 64), (3, 64, 64, 32), (1, 100, 72, 128),
                                        (1, 128, 128, 192)])
@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("tc", [True, False, "bulk-store"])
def test_stem_matches_fp64(n, h, w, cout, u8, tc):
    # tc = "bulk-store": the tcgen05 kernel with that drain forced (the default takes the bulk-store drain for
    # large batches only; csrc/conv_tc.cuh)
    drain = {"bulk-store": 5}.get(tc)
    if drain is not None:
        old = _lib.get_option("direct_epi")
        _lib.set_option("direct_epi", drain)
        try:
            return test_stem_matches_fp64(n, h, w, cout, u8, True)
        finally:
            _lib.set_option("direct_epi", old)
    conv = conv3x3(3, cout, 2)
    with torch.no_grad():
        conv.weight.copy_(uniform(tuple(conv.weight.shape), "stem.w", 1) / 27 ** 0.5)
        conv.bias.copy_(uniform(tuple(conv.bias.shape), "stem.b", 1) / 27 ** 0.5)
    conv = conv.cuda()
    x = uniform((n, 3, h, w), "stem.x", 2)
    if u8:
        xin = ((x + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
        xf = (xin.float() / 255.0 - 0.5) * 2                     # convert_image_dtype + (x - 0.5) * 2 (demo.py:110-118)
    else:
        xin = xf = x
    top, left, hp, wp = aligned_pad_amounts(h, w)
    xp = F.pad(xf, (left, wp - w - left, top, hp - h - top), "reflect") if (hp, wp) != (h, w) else xf
    exp = F.conv2d(xp.double(), conv.weight.detach().cpu().double(), conv.bias.detach().cpu().double(), stride=2,
                   padding=1).permute(0, 2, 3, 1)
    eng = Engine()
    eng.passes = 3
    eng.stem_tc = tc
    before = eng.lib.mcq_kernel_launch_count()
    for want in ({"f32", "silu"}, {"raw"}):
        out = eng.stem(conv, xin.cuda(), (top, left, hp, wp), want)
        torch.cuda.synchronize()
        scale = float(exp.abs().max())
        tol = 3e-6 if tc and cout <= 128 else 1e-6
        if "f32" in want:
            assert tuple(out.f32.shape) == (n, hp // 2, wp // 2, cout)
            assert float((out.f32.cpu().double() - exp).abs().max()) <= tol * scale
            assert float((_planes_value(out.silu).cpu() - F.silu(exp)).abs().max()) <= tol * scale
        else:
            assert float((_planes_value(out.raw).cpu() - exp).abs().max()) <= tol * scale
    assert eng.lib.mcq_kernel_launch_count() - before == 2
    assert eng.lib.mcq_device_error_flag() == 0


def test_unsupported_shapes_are_refused_by_the_tc_entry_point():
    eng = Engine()
    x = torch.zeros(1, 3, 64, 64, device="cuda")
    rows = torch.zeros(208, 64, dtype=torch.float16, device="cuda")
    bias = torch.zeros(200, device="cuda")
    out = torch.empty(1, 32, 32, 200, device="cuda")
    import ctypes
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = eng.lib.mcq_stem_conv_tc(p(x), 0, 1, 64, 64, 0, 0, 64, 64, p(rows), 1.0, p(bias), 200, 208, p(out), None, None, 0,
                                  None)
    assert rc == _lib.ERR_UNSUPPORTED
