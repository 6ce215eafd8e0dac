"""ResidualBlock / AttentionBlock variants of the `Neon` tokenizer: nn.GroupNorm in
place of the second SiLU (denseNorm=True, mcquic/nn/blocks.py:198) and channel-changing blocks with a conv1x1 skip
(:189-192).

CPU part: the oracle restatement is pinned bit-for-bit to the reference's own classes (where /root/reference exists)
and to the committed reference outputs tests/golden/blocks_dense.npz (made by oracle/gen_golden.py --blocks); the
product's host logic (fusion map: skip conv -> residual operand, GroupNorm launch between the convs) is run against
the CPU model of the C ABI.  GPU part: the CUDA path through the C ABI against the same golden outputs, the
mcq_groupnorm kernel alone against an fp64 evaluation, determinism, and error behaviour.
"""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import DENSE_BLOCKS, GOLDEN, dense_block_inputs, dense_block_oracle, dense_stride
from emulator import EmulatedLib
from mcquic_b200 import _lib
from mcquic_b200.engine import Act, Engine
from mcquic_b200.nn import blocks as B
from oracle import ref_import

# fp32 outputs of magnitude <= 5: 3-pass split-fp16 convs are fp32-grade (measured ~2e-6 relative per conv)
TOL_3PASS = 2e-5
# 1-pass (fp16 operands, TF32-grade): the budget north_star gives reconstructed pixels
TOL_1PASS = 1e-2


def _golden():
    return np.load(os.path.join(GOLDEN, "blocks_dense.npz"))


def _sample(y, shape):
    s = dense_stride(shape)
    return y[..., ::s, ::s]


@pytest.mark.parametrize("name", list(DENSE_BLOCKS))
def test_oracle_reproduces_reference_golden(name):
    kind, args, shape = DENSE_BLOCKS[name]
    block, x = dense_block_inputs(name, getattr(B, kind))
    y = dense_block_oracle(name, block.state_dict(), x)
    ref = torch.from_numpy(_golden()[name])
    assert tuple(_sample(y, shape).shape) == tuple(ref.shape)
    # same torch build => bit-identical; another oneDNN build may differ in the last ulps
    assert float((_sample(y, shape) - ref).abs().max()) <= 2e-6


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", list(DENSE_BLOCKS))
def test_oracle_bit_identical_to_reference_classes(name):
    ref_import.load()
    import mcquic.nn.blocks as RB
    kind, args, shape = DENSE_BLOCKS[name]
    ref_block, x = dense_block_inputs(name, getattr(RB, kind))
    mine, _ = dense_block_inputs(name, getattr(B, kind))
    assert list(ref_block.state_dict()) == list(mine.state_dict())       # same state_dict layout as the reference
    assert all(torch.equal(a, b) for a, b in zip(ref_block.state_dict().values(), mine.state_dict().values()))
    with torch.inference_mode():
        y = ref_block(x)
    assert torch.equal(y, dense_block_oracle(name, ref_block.state_dict(), x))
    assert torch.equal(_sample(y, shape), torch.from_numpy(_golden()[name]))


@pytest.mark.parametrize("name", [n for n in DENSE_BLOCKS if n != "rb128_gn1_64x64"])
def test_host_logic_through_emulated_abi(name):
    kind, args, shape = DENSE_BLOCKS[name]
    block, x = dense_block_inputs(name, getattr(B, kind))
    eng = Engine(lib=EmulatedLib())
    y = eng.run_module_nchw(block, x)
    ref = dense_block_oracle(name, block.state_dict(), x)
    assert float((y - ref).abs().max()) <= TOL_3PASS
    convs = sum(isinstance(m, torch.nn.Conv2d) for m in block.modules())
    norms = sum(isinstance(m, torch.nn.GroupNorm) for m in block.modules())
    assert eng.lib.launches == convs + norms + 2        # one launch per conv, one per GroupNorm, NCHW<->NHWC boundary


def test_needs_of_channel_changing_block():
    assert Engine.needs_of(B.ResidualBlock(64, 128)) == {"raw", "silu"}
    assert Engine.needs_of(B.ResidualBlock(64, 64, 8, True)) == {"f32", "silu"}
    assert isinstance(B.ResidualBlock(64, 64, 8, True)._branch[2], torch.nn.GroupNorm)
    assert isinstance(B.ResidualBlock(64, 64)._branch[2], torch.nn.SiLU)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(DENSE_BLOCKS))
@pytest.mark.parametrize("passes", [3, 1])
def test_gpu_blocks_against_reference_golden(name, passes):
    kind, args, shape = DENSE_BLOCKS[name]
    block, x = dense_block_inputs(name, getattr(B, kind))
    block = block.cuda()
    eng = Engine("tcgen05")
    eng.passes = passes
    before = _lib.launch_count()
    y = eng.run_module_nchw(block, x.cuda())
    torch.cuda.synchronize()
    assert eng.lib.mcq_device_error_flag() == 0
    assert _lib.launch_count() - before >= 4
    ref = torch.from_numpy(_golden()[name])
    err = float((_sample(y.cpu(), shape) - ref).abs().max())
    assert err <= (TOL_3PASS if passes == 3 else TOL_1PASS), err
    full = dense_block_oracle(name, {k: v.cpu() for k, v in block.state_dict().items()}, x)
    assert float((y.cpu() - full).abs().max()) <= (TOL_3PASS if passes == 3 else TOL_1PASS)
    assert torch.equal(y, eng.run_module_nchw(block, x.cuda()))          # fixed-order statistics: bit-reproducible


GN_CASES = [
    # n, h, w, c, groups
    (2, 64, 64, 128, 32),      # 8-CTA clusters
    (3, 16, 16, 128, 1),       # one group over everything, 8 slices of 32 pixels
    (5, 4, 4, 64, 8),          # tiny map: a single CTA per image
    (2, 7, 9, 192, 6),         # c/4 = 48 does not divide the block: idle lanes, ragged slices
    (1, 33, 1, 8, 8),          # one channel per group (instance norm), c/4 = 2
    (2, 128, 128, 256, 32),    # largest per-image working set
    (1, 3, 5, 512, 32),        # c at the limit
    (4, 10, 6, 32, 32),
]


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,c,groups", GN_CASES)
def test_gpu_groupnorm_kernel(n, h, w, c, groups):
    g = torch.Generator().manual_seed(n * 1000 + c)
    # a mean well away from zero (conv outputs have one): exercises the variance cancellation
    x = torch.randn(n, h, w, c, generator=g) * 0.7 + torch.randn(1, 1, 1, c, generator=g) * 1.5
    gamma = 1 + 0.3 * torch.randn(c, generator=g)
    beta = 0.2 * torch.randn(c, generator=g)
    norm = torch.nn.GroupNorm(groups, c)
    with torch.no_grad():
        norm.weight.copy_(gamma)
        norm.bias.copy_(beta)
    norm = norm.cuda()
    eng = Engine("tcgen05")
    eng.passes = 3
    act = Act(n, h, w, c, f32=x.cuda())
    exp = F.group_norm(x.double().permute(0, 3, 1, 2), groups, gamma.double(), beta.double(), norm.eps).permute(0, 2, 3, 1)
    scale = float(exp.abs().max())
    for want, fn in (({"f32", "raw"}, lambda v: v), ({"silu"}, F.silu), ({"f32"}, lambda v: v)):
        out = eng.groupnorm(norm, act, set(want))
        torch.cuda.synchronize()
        if "f32" in want:
            assert float((out.f32.cpu().double() - exp).abs().max()) <= 4e-6 * scale
            # and against PyTorch's own fp32 kernel (what the reference runs)
            ref32 = F.group_norm(x.permute(0, 3, 1, 2), groups, gamma, beta, norm.eps).permute(0, 2, 3, 1)
            assert float((out.f32.cpu() - ref32).abs().max()) <= 4e-6 * scale
        for name in want - {"f32"}:
            hi, lo = getattr(out, name)
            got = hi.cpu().double() + lo.cpu().double() / 2048.0
            assert float((got - fn(exp)).abs().max()) <= (2.0 ** -19 + 4e-6) * scale
    eng.passes = 1
    out = eng.groupnorm(norm, act, {"raw"})
    assert out.raw[1] is None
    assert float((out.raw[0].cpu().double() - exp).abs().max()) <= 2.0 ** -10 * scale
    assert eng.lib.mcq_device_error_flag() == 0


FUSED_CASES = [
    # n, h, w, c, groups, fused expected
    (2, 64, 64, 128, 32, True),      # unit 4 (four groups per 16-channel chunk)
    (3, 24, 20, 128, 1, True),       # one group, ragged tiles in both directions, unit 16
    (2, 16, 8, 256, 32, True),       # smallest map the pair kernel takes, two N tiles, unit 8
    (1, 40, 24, 128, 8, True),       # unit 16 = one group per chunk
    (5, 18, 30, 128, 4, True),       # rows not a multiple of the 4-row block
    (2, 16, 16, 128, 64, False),     # 2 channels per group: not expressible in channel quads -> stand-alone kernel
    (2, 8, 8, 128, 32, False),       # map below the pair kernel's tile -> stand-alone kernel
]


@pytest.mark.gpu
@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("n,h,w,c,groups,fused", FUSED_CASES)
def test_gpu_groupnorm_statistics_fused_into_the_conv_epilogue(n, h, w, c, groups, fused, passes):
    """conv3x3 whose epilogue also emits the GroupNorm partial sums (mcq_conv_params.gn_partials) + mcq_groupnorm_apply
    == the same conv followed by the stand-alone mcq_groupnorm == fp64 GroupNorm of the conv's fp32 output."""
    from convcase import make_planes
    from mcquic_b200.engine import pack_conv
    g = torch.Generator().manual_seed(h * 100 + c + groups)
    x = torch.randn(n, h, w, c, generator=g).cuda()
    wt = ((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
    bias = (torch.rand(c, generator=g) - 0.3).cuda()            # non-zero mean in the conv output
    norm = torch.nn.GroupNorm(groups, c)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(c, generator=g))
        norm.bias.copy_(0.2 * torch.randn(c, generator=g))
    norm = norm.cuda()
    eng = Engine("tcgen05")
    eng.passes = passes
    pc = pack_conv(wt, bias, 1, _lib.STORE_NHWC, "cuda")
    a = make_planes(x, passes)
    t = eng.conv(pc, a, Act(n, h, w, c), {"f32"}, gn_groups=groups)
    plain = eng.conv(pc, a, Act(n, h, w, c), {"f32"})
    eng.flush()
    torch.cuda.synchronize()
    assert (t.gn is not None) == fused
    assert torch.equal(t.f32, plain.f32)                        # the statistics do not disturb the conv's own output
    want = {"f32", "raw"}
    y_fused = eng.groupnorm(norm, t, want)
    y_alone = eng.groupnorm(norm, plain, want)
    torch.cuda.synchronize()
    exp = F.group_norm(t.f32.cpu().double().permute(0, 3, 1, 2), groups, norm.weight.detach().cpu().double(),
                       norm.bias.detach().cpu().double(), norm.eps).permute(0, 2, 3, 1)
    scale = float(exp.abs().max())
    for y in (y_fused, y_alone):
        assert float((y.f32.cpu().double() - exp).abs().max()) <= 4e-6 * scale
        if passes == 3:
            rec = y.raw[0].cpu().double() + y.raw[1].cpu().double() / 2048.0
            assert float((rec - exp).abs().max()) <= (2.0 ** -19 + 4e-6) * scale
        else:
            assert y.raw[1] is None and float((y.raw[0].cpu().double() - exp).abs().max()) <= 2.0 ** -10 * scale
    if fused:
        part, rb, unit = t.gn
        assert tuple(part.shape) == (n, rb, c // unit, 2) and rb == -(-h // 4) * -(-w // 8)
        # every slot written exactly once: the partials add up to the image sums
        tot = part.double().sum(1).reshape(n, c // unit, 2).cpu()
        ref = t.f32.cpu().double().reshape(n, h * w, c // unit, unit)
        assert float((tot[..., 0] - ref.sum((1, 3))).abs().max()) <= 1e-4 * float(ref.abs().sum((1, 3)).max())
        assert float((tot[..., 1] - (ref * ref).sum((1, 3))).abs().max()) <= 1e-5 * float((ref * ref).sum((1, 3)).max())
        # bit-reproducible (fixed-order reductions, no atomics)
        t2 = eng.conv(pc, a, Act(n, h, w, c), {"f32"}, gn_groups=groups)
        assert torch.equal(t2.gn[0], part) and torch.equal(eng.groupnorm(norm, t2, want).f32, y_fused.f32)
    assert eng.lib.mcq_device_error_flag() == 0


@pytest.mark.gpu
def test_gpu_groupnorm_errors():
    lib = _lib.load()
    x = torch.zeros(1, 4, 4, 6, device="cuda")
    gb = torch.zeros(1024, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.mcq_groupnorm(p(x), 1, 4, 4, 6, 3, p(gb), p(gb), 1e-5, p(x), None, None, 0, None) == _lib.ERR_UNSUPPORTED
    assert lib.mcq_groupnorm(p(x), 1, 4, 4, 8, 3, p(gb), p(gb), 1e-5, p(x), None, None, 0, None) == _lib.ERR_BAD_ARG
    assert lib.mcq_groupnorm(p(x), 1, 4, 4, 8, 2, p(gb), p(gb), 1e-5, None, None, None, 0, None) == _lib.ERR_BAD_ARG
    assert lib.mcq_groupnorm(p(x), 1, 1, 1, 1024, 2, p(gb), p(gb), 1e-5, p(x), None, None, 0, None) == _lib.ERR_UNSUPPORTED
    with pytest.raises(RuntimeError, match="CUDA"):
        B.ResidualBlock(64, 64, 8, True)(torch.zeros(1, 64, 8, 8))       # CPU tensor: no fallback


def test_groupnorm_bad_arguments_without_a_gpu():
    lib = _lib.load()
    assert lib.mcq_groupnorm(None, 1, 4, 4, 8, 2, None, None, 1e-5, None, None, None, 0, None) == _lib.ERR_BAD_ARG
