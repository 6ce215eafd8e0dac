"""Parity of the convolution kernels (tcgen05 and SIMT) through the C ABI against an fp64 torch evaluation of the
same operands.  Tolerances (relative to the largest expected magnitude):
  fp32 outputs, 3-pass split-fp16 or SIMT : 1e-5   (fp32-grade: measured ~2e-6 on B200)
  fp32 outputs, 1-pass                    : 1e-5   (vs the same hi-only operands; operand rounding is not kernel error)
  split-fp16 plane outputs                : 2^-19 with the lo plane (3-pass), 2^-10 hi only (1-pass)
"""
import pytest
import torch

from convcase import compare, run_case
from mcquic_b200 import _lib
from mcquic_b200.engine import Engine

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5


def _tol(name, passes):
    if name == "f32":
        return TOL_F32
    return 2.0 ** -19 + TOL_F32 if passes == 3 else 2.0 ** -10


def _check(impl, want=("f32",), **kw):
    eng = Engine(impl)
    out, exp, _ = run_case(eng, want=want, **kw)
    err = compare(out, exp, want, kw.get("passes", 3))
    assert eng.lib.mcq_device_error_flag() == 0
    for name, e in err.items():
        assert e <= _tol(name, kw.get("passes", 3)), (name, e, kw)


SHAPES = [
    dict(n=1, h=16, w=16, cin=128, cout=128),                 # one tile
    dict(n=4, h=64, w=64, cin=128, cout=128),                 # multi-tile, persistent loop, TMEM double buffering
    dict(n=4, h=8, w=8, cin=128, cout=128),                   # two images per tile
    dict(n=5, h=4, w=4, cin=128, cout=128),                   # eight images per tile, ragged batch
    dict(n=3, h=2, w=2, cin=128, cout=128),                   # 2x2 maps (128-pixel inputs, level 2)
    dict(n=2, h=24, w=40, cin=128, cout=128),                 # ragged H and W
    dict(n=3, h=6, w=12, cin=128, cout=128),                  # non power-of-two maps (384x768 inputs)
    dict(n=2, h=16, w=16, cin=192, cout=192),                 # C=192 (qp>=3 models): N tile 192
    dict(n=2, h=16, w=16, cin=64, cout=64),
    dict(n=2, h=32, w=32, cin=128, cout=128, stride=2),       # 5-D parity view
    dict(n=3, h=8, w=8, cin=128, cout=128, stride=2),
    dict(n=1, h=48, w=80, cin=128, cout=128, stride=2),
    dict(n=2, h=16, w=16, cin=128, cout=128, ksize=1),
]


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
def test_tcgen05_conv_shapes(shape, passes):
    _check("tcgen05", passes=passes, **shape)


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_epilogue_modes(impl):
    base = dict(n=2, h=16, w=16, cin=128, cout=128)
    _check(impl, want=("f32", "silu"), ksize=1, mode=_lib.EPI_GATE, **base)          # AttentionBlock tail
    _check(impl, want=("raw",), ksize=1, mode=_lib.EPI_GDN, **base)                   # GDN
    _check(impl, want=("raw",), ksize=1, mode=_lib.EPI_IGDN, **base)                  # inverse GDN
    _check(impl, want=("f32", "raw"), use_res1=True, res1_scale=-1.0, use_res2=True, **base)  # z - deq, q + side
    _check(impl, want=("silu", "sq"), passes=1, **base)
    _check(impl, want=("f32", "sq"), n=2, h=16, w=16, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC)
    _check(impl, n=2, h=32, w=32, cin=128, cout=12, store=_lib.STORE_SHUFFLE_NCHW)   # last layer, NCHW pixels
    _check(impl, n=2, h=32, w=32, cin=128, cout=12, store=_lib.STORE_SHUFFLE_NCHW, passes=1)


def test_weight_stationary_pair_kernel():
    """1-pass 3x3 convs with cin = 128 run on the CTA-pair kernel with the weights resident in shared memory
    (conv_pair.cuh): many pixel tiles per cluster (the weights are loaded for the first one only), ragged maps, and the
    512-column PixelShuffle convs where every cluster serves one of the four N tiles."""
    _check("tcgen05", passes=1, want=("f32", "silu"), use_res1=True, n=24, h=64, w=64, cin=128, cout=128)
    _check("tcgen05", passes=1, want=("silu",), n=3, h=40, w=24, cin=128, cout=128)
    _check("tcgen05", passes=1, want=("f32", "sq"), n=20, h=32, w=32, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC)
    _check("tcgen05", passes=1, want=("f32",), n=1, h=16, w=16, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC)
    _check("tcgen05", passes=1, want=("raw",), n=2, h=16, w=24, cin=128, cout=256)


@pytest.fixture
def drain_option(request):
    old = _lib.get_option("direct_epi")
    _lib.set_option("direct_epi", request.param)
    yield request.param
    _lib.set_option("direct_epi", old)


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("drain_option", [1, 5], indirect=True, ids=["rows", "bulk-store-everywhere"])
def test_drain_variants(drain_option, passes):
    """both drains (csrc/conv_tc.cuh: row-per-lane global stores, bulk-tensor stores through shared memory),
    forced onto every launch that can take it -- the default only picks the bulk-store drain for large launches: ragged
    and tiny maps (boxes clipped at image borders, boxes spanning images), PixelShuffle stores, stride 2, every epilogue
    mode, two plane pairs, both residual operands"""
    base = dict(n=2, h=16, w=16, cin=128, cout=128, passes=passes)
    _check("tcgen05", want=("f32", "silu"), use_res1=True, n=24, h=64, w=64, cin=128, cout=128, passes=passes)
    _check("tcgen05", want=("f32", "silu"), use_res1=True, n=3, h=40, w=24, cin=128, cout=128, passes=passes)   # ragged
    _check("tcgen05", want=("silu",), n=2, h=24, w=40, cin=128, cout=128, passes=passes)
    _check("tcgen05", want=("f32", "raw"), use_res1=True, n=5, h=4, w=4, cin=128, cout=128, passes=passes)     # 8 images / tile
    _check("tcgen05", want=("f32",), n=3, h=2, w=2, cin=128, cout=128, passes=passes)
    _check("tcgen05", want=("f32", "silu"), n=3, h=6, w=12, cin=128, cout=128, passes=passes)
    _check("tcgen05", want=("f32", "sq"), n=2, h=32, w=32, cin=128, cout=128, stride=2, passes=passes)
    _check("tcgen05", want=("f32", "sq"), n=20, h=32, w=32, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC, passes=passes)
    _check("tcgen05", want=("f32", "sq"), n=1, h=10, w=12, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC, passes=passes)
    _check("tcgen05", want=("f32", "silu"), ksize=1, mode=_lib.EPI_GATE, **base)
    _check("tcgen05", want=("raw",), ksize=1, mode=_lib.EPI_GDN, **base)
    _check("tcgen05", want=("raw",), ksize=1, mode=_lib.EPI_IGDN, **base)
    _check("tcgen05", want=("f32", "raw"), use_res1=True, res1_scale=-1.0, use_res2=True, **base)
    _check("tcgen05", want=("silu", "sq"), **base)
    _check("tcgen05", want=("f32",), n=2, h=16, w=16, cin=192, cout=192, passes=passes)        # halo kernel, N tile 192
    _check("tcgen05", want=("raw",), n=2, h=16, w=24, cin=128, cout=256, passes=passes)


PARTIAL_CHUNK_SHAPES = [
    # channel counts the 64-channel K chunk does not divide (Neon's 8 / 32-channel nets): the last chunk is zero-filled
    # by TMA beyond the tensor's channel extent; stride 1 only
    dict(n=2, h=64, w=64, cin=32, cout=32),                  # halo kernel, 2 tiles per image column
    dict(n=3, h=24, w=20, cin=32, cout=32),                  # ragged
    dict(n=2, h=16, w=16, cin=32, cout=128),                 # CTA-pair kernel (3-pass) / weight... 32 -> 128
    dict(n=2, h=16, w=16, cin=32, cout=128, store=_lib.STORE_SHUFFLE_NHWC),
    dict(n=4, h=4, w=4, cin=32, cout=32),                    # per-tap kernel, several images per tile
    dict(n=2, h=8, w=8, cin=8, cout=32),                     # 8 of 64 channels real
    dict(n=2, h=32, w=16, cin=32, cout=8),                   # 32 -> 8 (N tile 16)
    dict(n=2, h=16, w=16, cin=96, cout=64),                  # 1.5 chunks
    dict(n=2, h=16, w=16, cin=32, cout=32, ksize=1),
    dict(n=1, h=12, w=40, cin=40, cout=64, ksize=1),
]


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("shape", PARTIAL_CHUNK_SHAPES, ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
def test_tcgen05_conv_partial_k_chunk(shape, passes):
    from mcquic_b200.engine import Engine as E
    before = _lib.launch_count()
    _check("tcgen05", passes=passes, **shape)
    assert _lib.launch_count() > before


def test_simt_serves_channel_counts_the_tensor_core_tiling_does_not():
    _check("simt", n=1, h=9, w=7, cin=32, cout=40)
    _check("simt", n=2, h=16, w=24, cin=64, cout=64, stride=2)
    _check("tcgen05", n=2, h=16, w=24, cin=32, cout=32, stride=2)      # engine routes stride 2 at cin % 64 != 0 to SIMT
    _check("tcgen05", n=1, h=9, w=7, cin=36, cout=40)                   # cin % 8 != 0: SIMT
    eng = Engine("tcgen05")
    with pytest.raises(RuntimeError, match="not supported|bad argument"):
        p = _lib.ConvParams()
        x = torch.zeros(1, 8, 8, 36, dtype=torch.float16, device="cuda")
        w = torch.zeros(32, 9 * 36, dtype=torch.float16, device="cuda")
        b = torch.zeros(32, device="cuda")
        o = torch.zeros(1, 8, 8, 32, device="cuda")
        p.a_hi = p.a_lo = x.data_ptr(); p.w_hi = p.w_lo = w.data_ptr(); p.bias = b.data_ptr(); p.out_f32 = o.data_ptr()
        p.n, p.hin, p.win, p.cin, p.cout, p.cout_pad, p.ksize, p.stride, p.passes = 1, 8, 8, 36, 32, 32, 3, 1, 3
        p.w_scale = 1.0
        p.impl = _lib.IMPL_TCGEN05
        import ctypes
        _lib.check(eng.lib.mcq_conv2d(ctypes.byref(p), None), "mcq_conv2d")


def test_tcgen05_equals_simt_on_a_full_size_layer():
    """same operands, two independent kernels (tensor cores vs fp32 FFMA): agree to fp32 rounding at N=64, 64x64"""
    kw = dict(n=64, h=64, w=64, cin=128, cout=128, passes=3, want=("f32",))
    a, _, _ = run_case(Engine("tcgen05"), **kw)
    b, _, _ = run_case(Engine("simt"), **kw)
    scale = float(b.f32.abs().max())
    assert float((a.f32 - b.f32).abs().max()) / scale <= TOL_F32


def test_linearity_at_full_size():
    """size-independent property: conv(x1 + x2) - bias == (conv(x1) - bias) + (conv(x2) - bias) up to rounding"""
    from convcase import make_planes
    from mcquic_b200.engine import Act, pack_conv
    g = torch.Generator().manual_seed(5)
    n, h, w, c = 64, 32, 32, 128
    x1 = torch.randn(n, h, w, c, generator=g).cuda()
    x2 = torch.randn(n, h, w, c, generator=g).cuda()
    wt = ((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
    pc = pack_conv(wt, torch.zeros(c).cuda(), 1, 0, "cuda")
    eng = Engine("tcgen05")
    eng.passes = 3
    f = lambda x: eng.conv(pc, make_planes(x, 3), Act(n, h, w, c), {"f32"}).f32
    lhs, rhs = f(x1 + x2), f(x1) + f(x2)
    assert float((lhs - rhs).abs().max()) <= 2e-5 * float(rhs.abs().max())
