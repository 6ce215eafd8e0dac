"""CPU oracle for McQuic's `Compressor.encode/decode` hot path  --  TEST INFRASTRUCTURE ONLY.

A functional (no nn.Module) fp32 restatement of the reference algorithm, operating directly on a
reference-layout ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file; the product (``mcquic_b200``)
never does and fails loudly when its CUDA library is missing.

Parity pin: ``tests/test_oracle_vs_reference.py`` runs this restatement against the *unmodified*
reference sources imported from /root/reference (``oracle/ref_import.py``) and requires bit-identical
codes and pixels on CPU; ``tests/golden/*.npz`` (made by ``oracle/gen_golden.py`` from the reference
itself) pin it on machines where the reference tree is absent (the GPU box).  The reference ships
no golden vectors or known-answer tests of its own for this path.

Arithmetic that lives outside /root/reference: conv/GEMM kernels are PyTorch's (oneDNN / MKL on CPU,
``torch>2`` unpinned in the reference's setup.py:48; here torch 2.11.0).  True fp32 everywhere.

Every function cites the reference lines (relative to /root/reference) it restates.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

ALIGN_BASE = 128  # mcquic/data/transforms.py:82  AlignedPadding(base=128)


# ----------------------------------------------------------------------------------------------
# mcquic/data/transforms.py:86-99
def aligned_padding(x: torch.Tensor, base: int = ALIGN_BASE) -> torch.Tensor:
    """Reflect-pad H and W up to the next multiple of `base`; the smaller half goes left/top."""
    h, w = x.shape[-2], x.shape[-1]
    wp = ((w // base + 1) * base - w) % base
    hp = ((h // base + 1) * base - h) % base
    left, top = wp // 2, hp // 2
    if wp == 0 and hp == 0:
        return x
    return F.pad(x, (left, wp - left, top, hp - top), "reflect")


# ----------------------------------------------------------------------------------------------
# mcquic/nn/convs.py:77-100 (conv3x3), :257-276 (conv1x1)
def _conv(sd: StateDict, p: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], stride=stride, padding=w.shape[-1] // 2)


# mcquic/nn/convs.py:221-255 -- conv(C -> r*r*Cout, 3x3) followed by nn.PixelShuffle(r)
def _pixel_shuffle_conv(sd: StateDict, p: str, x: torch.Tensor, r: int = 2) -> torch.Tensor:
    return F.pixel_shuffle(_conv(sd, p + ".0", x), r)


# mcquic/nn/base.py:58-84 NonNegativeParametrizer.forward + :31-54 LowerBound
def _reparam(sd: StateDict, p: str, raw: torch.Tensor) -> torch.Tensor:
    bound = sd[p + ".lowerBound.bound"]
    return torch.max(raw, bound) ** 2 - sd[p + ".eps"]


# mcquic/nn/gdn.py:67-91
def _gdn(sd: StateDict, p: str, x: torch.Tensor, inverse: bool) -> torch.Tensor:
    beta = _reparam(sd, p + ".beta_reparam", sd[p + ".beta"])
    gamma = _reparam(sd, p + ".gamma_reparam", sd[p + ".gamma"])
    norm = F.conv2d(x ** 2, gamma[..., None, None], beta)
    return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


# mcquic/nn/blocks.py:62-78 (_residulBlock.forward) + :163-200 (ResidualBlock)
# groups: nn.GroupNorm(groups, C) replaces the second SiLU when the block was built with denseNorm=True (:198; then
# there is NO activation before the second conv) -- recognised by the norm's `_branch.2.weight` in the state_dict;
# a `_skip` conv1x1 exists iff the channel count changes (:189-192) and takes the un-activated x.
def residual_block(sd: StateDict, p: str, x: torch.Tensor, groups: int = 1, eps: float = 1e-5) -> torch.Tensor:
    out = _conv(sd, p + "._branch.1", F.silu(x))
    if p + "._branch.2.weight" in sd:
        out = F.group_norm(out, groups, sd[p + "._branch.2.weight"], sd[p + "._branch.2.bias"], eps)
    else:
        out = F.silu(out)
    out = _conv(sd, p + "._branch.3", out)
    identity = _conv(sd, p + "._skip", x) if p + "._skip.weight" in sd else x
    return out + identity


# mcquic/nn/blocks.py:82-122 (ResidualBlockWithStride): SiLU, conv3 s2, GDN, conv3; skip = conv3 s2 on raw x
def residual_block_stride(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    out = _conv(sd, p + "._branch.1", F.silu(x), stride=2)
    out = _gdn(sd, p + "._branch.2", out, inverse=False)
    out = _conv(sd, p + "._branch.3", out)
    return out + _conv(sd, p + "._skip", x, stride=2)


# mcquic/nn/blocks.py:125-159 (ResidualBlockShuffle): SiLU, pixShuf3, IGDN, conv3; skip = pixShuf3 on raw x
def residual_block_shuffle(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    out = _pixel_shuffle_conv(sd, p + "._branch.1", F.silu(x))
    out = _gdn(sd, p + "._branch.2", out, inverse=True)
    out = _conv(sd, p + "._branch.3", out)
    return out + _pixel_shuffle_conv(sd, p + "._skip", x)


# mcquic/nn/blocks.py:246-288 (AttentionBlock); groups / denseNorm are handed to its six ResidualBlocks (:264-273)
def attention_block(sd: StateDict, p: str, x: torch.Tensor, groups: int = 1) -> torch.Tensor:
    a = x
    for i in range(3):
        a = residual_block(sd, f"{p}._mainBranch.{i}", a, groups)
    b = x
    for i in range(3):
        b = residual_block(sd, f"{p}._sideBranch.{i}", b, groups)
    b = _conv(sd, p + "._sideBranch.3", b)
    return a * torch.sigmoid(b) + x


_RB, _RBS, _RBU, _AB, _CONV = residual_block, residual_block_stride, residual_block_shuffle, attention_block, _conv


def _run_from(sd: StateDict, p: str, start: int, blocks: Sequence, x: torch.Tensor) -> torch.Tensor:
    for i, fn in enumerate(blocks):
        x = fn(sd, f"{p}.{start + i}", x)
    return x


# mcquic/modules/compressor.py:122-131 (analysis transform), :132-140 (synthesis transform)
def analysis(sd: StateDict, x: torch.Tensor) -> torch.Tensor:
    x = _conv(sd, "_encoder.0", x, stride=2)
    return _run_from(sd, "_encoder", 1, [_RB, _RBS, _AB, _RB, _RBS, _RB], x)


def synthesis(sd: StateDict, y: torch.Tensor) -> torch.Tensor:
    y = _run_from(sd, "_decoder", 0, [_RB, _RBU, _AB, _RB, _RBU, _RB], y)
    return _pixel_shuffle_conv(sd, "_decoder.6", y)


# ----------------------------------------------------------------------------------------------
# mcquic/modules/quantizer.py:153-179 (_distance) -- returns [n, m, h, w, k]
def vq_distance(x: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    n, _, h, w = x.shape
    m, k, d = codebook.shape
    x = x.reshape(n, m, d, h, w)
    x2 = (x ** 2).sum(2, keepdim=True)                                  # [n, m, 1, h, w]
    c2 = (codebook ** 2).sum(-1, keepdim=True)[..., None]               # [m, k, 1, 1]
    left = x.reshape(n * m, d, h * w).permute(0, 2, 1).contiguous()     # [nm, hw, d]
    right = codebook.expand(n, m, k, d).reshape(n * m, k, d).permute(0, 2, 1).contiguous()
    inter = torch.bmm(left, right).reshape(n, m, h, w, k).permute(0, 1, 4, 2, 3)
    distance = x2 + c2 - 2 * inter                                      # [n, m, k, h, w]
    return distance.permute(0, 1, 3, 4, 2).contiguous()


# mcquic/modules/quantizer.py:144-150 (encode): argmin over k, first index on ties
def vq_assign(x: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    return vq_distance(x, codebook).argmin(-1)


# mcquic/modules/quantizer.py:181-183 (_logit) and :204 (x LowerBound(Eps)(temperature)); deterministic part
# of the training-time soft path (before _randomDrop / gumbelSoftmax, which consume RNG).
def vq_logits(x: torch.Tensor, codebook: torch.Tensor, temperature: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    k = codebook.shape[1]
    logit = (-1 * vq_distance(x, codebook)) / (k ** 0.5)
    return logit * torch.max(temperature, torch.tensor([eps]))


# mcquic/modules/quantizer.py:249-259 (_multiCodebookDeQuantization.decode)
def vq_dequantize(code: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    n, m, h, w = code.shape
    ix = torch.arange(m)[None, None, None, :].expand(n, h, w, m)
    picked = codebook[ix, code.permute(0, 2, 3, 1)]                     # [n, h, w, m, d]
    return picked.reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


def vq_margin(x: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """Relative gap between best and second-best distance per code, [n, m, h, w] (diagnostic:
    a GPU/CPU index flip at margin < ~1e-6 is rounding, not a bug -- SURVEY.md section 7)."""
    top2 = torch.topk(vq_distance(x, codebook), 2, dim=-1, largest=False).values
    return (top2[..., 1] - top2[..., 0]) / top2[..., 1].abs().clamp_min(1e-30)


# ----------------------------------------------------------------------------------------------
def _levels(sd: StateDict) -> int:
    lv = 0
    while f"_quantizer._encoders.{lv}._quantizer._codebook" in sd:
        lv += 1
    return lv


# mcquic/modules/compressor.py:142-154 per-level nets; mcquic/modules/quantizer.py:310-318, :411-420
def quantizer_encode(sd: StateDict, y: torch.Tensor, with_margin: bool = False):
    codes, margins = [], []
    levels = _levels(sd)
    x = y
    for lv in range(levels):
        p = f"_quantizer._encoders.{lv}"
        cb = sd[p + "._quantizer._codebook"]
        z = _run_from(sd, p + "._latentStageEncoder", 0, [_RBS, _RB, _AB], x)
        head = _run_from(sd, p + "._quantizationHead", 0, [_RB, _AB, _CONV], z)
        code = vq_assign(head, cb)
        codes.append(code)
        if with_margin:
            margins.append(vq_margin(head, cb))
        if lv < levels - 1:
            z = _run_from(sd, p + "._latentHead", 0, [_RB, _AB, _CONV], z)
            x = z - vq_dequantize(code, cb)
    return (codes, margins) if with_margin else codes


# mcquic/modules/compressor.py:161-175 per-level nets; mcquic/modules/quantizer.py:351-357, :422-428
def quantizer_decode(sd: StateDict, codes: List[torch.Tensor]) -> torch.Tensor:
    levels = len(codes)
    former: Optional[torch.Tensor] = None
    for lv in reversed(range(levels)):
        p = f"_quantizer._decoders.{lv}"
        cb = sd[p + "._dequantizer._codebook"]
        q = _run_from(sd, p + "._dequantizationHead", 0, [_AB, _CONV, _RB], vq_dequantize(codes[lv], cb))
        if lv < levels - 1:
            q = q + _run_from(sd, p + "._sideHead", 0, [_AB, _CONV, _RB], former)
        former = _run_from(sd, p + "._restoreHead", 0, [_AB, _RB, _RBU], q)
    return former


# mcquic/modules/compressor.py:79-88
@torch.inference_mode()
def encode(sd: StateDict, x: torch.Tensor, with_margin: bool = False):
    return quantizer_encode(sd, analysis(sd, aligned_padding(x)), with_margin)


# mcquic/modules/compressor.py:114-117
@torch.inference_mode()
def decode(sd: StateDict, codes: List[torch.Tensor]) -> torch.Tensor:
    return synthesis(sd, quantizer_decode(sd, codes))


# ----------------------------------------------------------------------------------------------
# `Neon` tokenizer: mcquic/modules/compressor.py:181-233 + ResidualBackwardQuantizer,
# mcquic/modules/quantizer.py:577-700.  Every ResidualBlock / AttentionBlock carries (groups, denseNorm); GroupNorm
# groups are 32 in the trunk and 1 inside the quantizer and for the blocks that touch its 8 channels.
NEON_LATENT = 8  # ResidualBackwardQuantizer.channel, quantizer.py:583


def _rb(groups):
    return lambda sd, p, x: residual_block(sd, p, x, groups)


def _ab(groups):
    return lambda sd, p, x: attention_block(sd, p, x, groups)


def _conv_nobias(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:   # conv1x1(..., bias=False), convs.py:257-276
    return F.conv2d(x, sd[p + ".weight"], None)


# compressor.py:186-206
def neon_analysis(sd: StateDict, x: torch.Tensor) -> torch.Tensor:
    x = _conv(sd, "_encoder.0", x)
    blocks = [_ab(32), _rb(32), _rb(32), _RBS, _rb(32), _RBS, _rb(32), _RBS, _ab(32), _rb(32), _rb(32), _rb(32), _rb(32),
              _rb(1), _ab(1)]
    return _run_from(sd, "_encoder", 1, blocks, x)


# compressor.py:207-227
def neon_synthesis(sd: StateDict, y: torch.Tensor) -> torch.Tensor:
    blocks = [_ab(1), _rb(1), _rb(32), _rb(32), _rb(32), _rb(32), _ab(32), _rb(32), _RBU, _rb(32), _RBU, _rb(32), _RBU,
              _rb(32), _rb(32), _ab(32)]
    y = _run_from(sd, "_decoder", 0, blocks, y)
    return _conv(sd, "_decoder.16", y)


def _neon_strided(size: Sequence[int]) -> List[bool]:
    """per level: does its stage halve the resolution (quantizer.py:600-657: thisSize == lastSize // 2)"""
    out, last = [], size[0] * 2
    for this in size:
        if this not in (last, last // 2):
            raise ValueError("The given size sequence does not half or equal to from left to right.")
        out.append(this == last // 2)
        last = this
    return out


# quantizer.py:600-657 -- the three per-level nets; the middle block is strided / shuffling or a plain ResidualBlock
def _neon_stage(sd, p, x, strided):
    return _conv_nobias(sd, p + ".3", _run_from(sd, p, 0, [_rb(1), _ab(1), _RBS if strided else _rb(1)], x))


def _neon_up(sd, p, x, strided):
    x = _conv_nobias(sd, p + ".0", x)
    return _run_from(sd, p, 1, [_RBU if strided else _rb(1), _ab(1), _rb(1)], x)


# quantizer.py:675-693 (ResidualBackwardQuantizer.encode): all latents first, then residual codes from the smallest
# level back to the largest; codes are returned smallest level first
def neon_quantizer_encode(sd: StateDict, y: torch.Tensor, size: Sequence[int], with_margin: bool = False):
    strided = _neon_strided(size)
    latents, x = [], y
    for lv in range(len(size)):
        x = _neon_stage(sd, f"_quantizer._encoders.{lv}", x, strided[lv])
        latents.append(x)
    codes, margins = [], []
    current = torch.zeros_like(latents[-1])
    for lv in reversed(range(len(size))):
        cb = sd[f"_quantizer._quantizers.{lv}._codebook"]
        residual = latents[lv] - current
        code = vq_assign(residual, cb)
        codes.append(code)
        if with_margin:
            margins.append(vq_margin(residual, cb))
        quantized = vq_dequantize(code, cb)
        current = quantized if lv == len(size) - 1 else _neon_up(sd, f"_quantizer._backwards.{lv}", quantized, strided[lv])
    return (codes, margins) if with_margin else codes


# quantizer.py:695-703 (ResidualBackwardQuantizer.decode); codes smallest level first
def neon_quantizer_decode(sd: StateDict, codes: List[torch.Tensor], size: Sequence[int]) -> torch.Tensor:
    strided = _neon_strided(size)
    former: Optional[torch.Tensor] = None
    for lv, code in zip(reversed(range(len(size))), codes):
        q = vq_dequantize(code, sd[f"_quantizer._dequantizers.{lv}._codebook"])
        former = _neon_up(sd, f"_quantizer._decoders.{lv}", q if former is None else q + former, strided[lv])
    return former


@torch.inference_mode()
def neon_encode(sd: StateDict, x: torch.Tensor, size: Sequence[int], with_margin: bool = False):
    return neon_quantizer_encode(sd, neon_analysis(sd, aligned_padding(x)), size, with_margin)


@torch.inference_mode()
def neon_decode(sd: StateDict, codes: List[torch.Tensor], size: Sequence[int]) -> torch.Tensor:
    return neon_synthesis(sd, neon_quantizer_decode(sd, codes, size))


# ----------------------------------------------------------------------------------------------
# mcquic/validate/handlers.py:138-172 (IdealBPP: torch.bincount per level/codebook) and the counting half of
# mcquic/modules/entropyCoder.py:28-36 (one-hot .sum((0,2,3)) == per-(m, k) occurrence count).
def code_histogram(codes: List[torch.Tensor], ks: Sequence[int]) -> List[torch.Tensor]:
    out = []
    for code, k in zip(codes, ks):
        m = code.shape[1]
        out.append(torch.stack([torch.bincount(code[:, j].flatten(), minlength=k) for j in range(m)]))
    return out


# mcquic/utils/vision.py:135-146 (DeTransform: [-1,1] float -> uint8, truncating)
def to_uint8(x: torch.Tensor) -> torch.Tensor:
    x = (x - (-1.0)) / (1.0 - (-1.0))
    return (x * (255 + 1.0 - 1e-3)).clamp(0.0, 255.0).byte()


# mcquic/validate/metrics.py:264-274 (PSNR.forward, upperBound 255, +1e-4 in the denominator)
def psnr_uint8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mse = ((a.double() - b.double()) ** 2).mean(dim=(1, 2, 3))
    return 10.0 * (255.0 ** 2 / (mse + 1e-4)).log10()
