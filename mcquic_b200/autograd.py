"""Training step on the CUDA path (SURVEY.md section 8f NEXT-3): differentiable forward of the reference's blocks
(mcquic/modules/compressor.py:35-43 -> mcquic/nn/blocks.py, mcquic/nn/gdn.py, mcquic/modules/quantizer.py:181-274,727-765).

Every convolution -- >99 % of the FLOPs of the training step -- runs forward AND backward in libmcquic_b200.so:

  forward   mcq_split_planes (fp32 NHWC -> fp16 operand plane) + mcq_conv2d (tcgen05, one fp16 pass, fp32 accumulation:
            TF32-grade, which is what the reference trains with -- mcquic/train/utils.py turns allow_tf32 on)
  dgrad     mcq_conv2d on the output gradient with transformed weights: 3x3 stride 1 -> taps flipped, channels
            transposed; 1x1 -> transposed; 3x3 stride 2 -> its sub-pixel form (four 3x3 kernels with the taps each output
            parity receives, MCQ_STORE_SHUFFLE_NHWC puts the four results where they belong)
  wgrad     mcq_conv_wgrad (csrc/conv_wgrad.cuh): pixels are the reduction axis, both operands the NHWC planes above
            consumed as MN-major tcgen05 operands, split-K over pixel tiles with a deterministic reduction
  gradients are brought into fp16's range by a power-of-two factor computed ON THE DEVICE per convolution (amax of the
  incoming gradient; no host sync) and removed by the epilogue / the wgrad reduction through a device scalar.

Tensors between the convolutions are ordinary fp32 torch tensors in channels_last memory (= the NHWC the kernels use, no
layout pass); the element-wise parts of a block (SiLU, GroupNorm, the GDN arithmetic, sigmoid gate, residual adds,
PixelShuffle) and the soft quantizer's softmax / straight-through estimator are torch operations whose backward autograd
derives -- stated plainly: in this round only the convolutions of the training step are hand-written kernels.
"""
import ctypes
import weakref
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .engine import Act, Engine, pack_conv, _ptr

_ENGINE: Optional[Engine] = None
# tensor-core passes of the forward and dgrad convolutions: 1 = one fp16 pass (TF32-grade, the default: what the reference
# trains with), 3 = split-fp16 x3 (fp32-grade; the parity tests use it so that the sampled codes -- hard argmax decisions --
# equal the fp32 reference's).  The weight gradient always runs one pass.
PASSES = 1


def set_passes(passes: int):
    global PASSES
    if passes not in (1, 3):
        raise ValueError("passes must be 1 or 3")
    PASSES = passes


def train_engine() -> Engine:
    """the engine the training-step convolutions launch through: single stream (autograd orders the work)"""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine()
        _ENGINE.multistream = False
        _ENGINE.chain = False
    _ENGINE.passes = PASSES
    return _ENGINE


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:
        return 0


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """[n, c, h, w] tensor -> the same values with NHWC memory (channels_last), fp32"""
    return x.float().contiguous(memory_format=torch.channels_last)


def _split(eng: Engine, x_nhwc_mem: torch.Tensor, n: int, h: int, w: int, c: int, pad_to: int = 8,
           dev_scale: Optional[torch.Tensor] = None):
    """fp32 tensor whose memory is [n, h, w, c] -> fp16 planes (hi, lo or None) [n, h, w, c_pad] of x * (*dev_scale)"""
    if c % pad_to != 0:      # RGB input / 3-channel output gradient: zero channels up to the vector width of the kernels
        cp = (c + pad_to - 1) // pad_to * pad_to
        padded = x_nhwc_mem.new_zeros((n, cp, h, w)).contiguous(memory_format=torch.channels_last)
        padded[:, :c] = x_nhwc_mem
        x_nhwc_mem, c = padded, cp
    hi = torch.empty((n, h, w, c), dtype=torch.float16, device=x_nhwc_mem.device)
    lo = torch.empty_like(hi) if eng.passes == 3 else None
    _lib.check(eng.lib.mcq_split_planes(_ptr(x_nhwc_mem), x_nhwc_mem.numel(), _lib.ACT_NONE, _ptr(hi), _ptr(lo),
                                        _ptr(dev_scale), eng._stream()), "mcq_split_planes")
    return hi, lo


def _grad_scale(g: torch.Tensor):
    """(S, 1/S) as device scalars: S = the power of two that brings max |g| to [256, 512) -- fp16 planes then keep ~2^-23
    of the largest gradient before flushing.  A handful of tiny launches, no host synchronisation."""
    amax = torch.linalg.vector_norm(g.detach(), float("inf")).float()       # one reduction launch, no |g| temporary
    s = torch.exp2(torch.floor(torch.log2(256.0 / amax.clamp_min(1e-37))).clamp(-100.0, 100.0))
    s = torch.where(torch.isfinite(s) & (amax > 0), s, torch.ones_like(s)).reshape(1).contiguous()
    return s, (1.0 / s).contiguous()


def _dgrad_weight(weight: torch.Tensor, stride: int) -> torch.Tensor:
    """nn.Conv2d weight [cout, cin, k, k] -> the weight of the convolution that maps dY to dX (see module docstring)"""
    cout, cin, k, _ = weight.shape
    w = weight.detach().float()
    if stride == 1:
        return w.flip(2, 3).transpose(0, 1).contiguous() if k == 3 else w.transpose(0, 1).contiguous()
    # stride 2, 3x3, padding 1: dX[2y+i, 2x+j] gathers dY[y + dy] with forward tap r = 2 * (-dy) + i + 1 (same for columns)
    out = w.new_zeros((4 * cin, cout, 3, 3))              # rows in PixelShuffle order 4 * ci + 2 * i + j
    taps = {0: [(1, 1)], 1: [(1, 2), (2, 0)]}             # parity -> [(tap index of the dY conv, forward tap r)]
    wt = w.transpose(0, 1)                                # [cin, cout, r, s]
    for i in (0, 1):
        for ty, r in taps[i]:
            for j in (0, 1):
                for tx, s in taps[j]:
                    out[2 * i + j::4, :, ty, tx] = wt[:, :, r, s]
    return out


class _Packed:
    """per-layer cache of the forward / dgrad operand packings.  The packings are valid for one weight version; the
    power-of-two exponents they were scaled with are kept for the life of the layer: re-packing after an optimizer step
    is then a handful of device operations without any host read, so a whole training step (forward, backward, optimizer)
    can be captured in ONE CUDA graph.  (fp16 operands scaled to max |w| = 256..512 have a factor 128 of head-room and 2^-22
    of resolution below that, far more than weights drift between re-scalings; `reset_weight_scales()` re-derives them.)"""

    def __init__(self, owner=None):
        self.owner = None if owner is None else weakref.ref(owner)   # id(owner) is reused once the module is collected
        self.key = None
        self.fwd = None
        self.dgrad = None
        self.exp_fwd = None
        self.exp_dgrad = None


_PACKS = {}


def reset_weight_scales():
    """forget the cached packings and exponents (call outside a captured step, e.g. every few thousand steps)"""
    _PACKS.clear()


def _exp_of(pc) -> int:
    import math
    return int(round(-math.log2(pc.w_scale)))


def _packs_for(owner: nn.Module, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> _Packed:
    pk = _PACKS.get(id(owner))
    key = (_version(weight), weight.data_ptr(), tuple(weight.shape), None if bias is None else _version(bias))
    capturing = weight.is_cuda and torch.cuda.is_current_stream_capturing()
    if pk is None or pk.owner is None or pk.owner() is not owner:
        # (a collected module's id -- and, through the caching allocator, even its weight's address and version -- can come
        #  back with another layer: the entry belongs to the module it was made for, nothing else)
        pk = _PACKS[id(owner)] = _Packed(owner)
    if pk.key != key or capturing or weight.grad_fn is not None:
        # new weight values (or a graph capture, whose replays must re-pack, or weights that are themselves computed each
        # step like GDN's reparametrised gamma): drop the packings, keep the exponents
        pk.key, pk.fwd, pk.dgrad = key, None, None
    return pk


def conv_supported(conv: nn.Conv2d) -> bool:
    """shapes the training-step kernels take; anything else (the 32-channel stride-2 convolutions of Neon's quantizer
    nets: a stride-2 tap view needs 64-channel K chunks) goes through torch's convolution"""
    k, s = conv.kernel_size[0], conv.stride[0]
    if conv.groups != 1 or conv.dilation[0] != 1 or k not in (1, 3) or s not in (1, 2) or conv.padding[0] != k // 2:
        return False
    if s == 2 and (conv.in_channels % 64 != 0 or k != 3 or conv.out_channels % 8 != 0):
        return False
    # channel counts the 16-byte vector accesses address: multiples of 8, or RGB (zero-padded to 8)
    return all(c % 8 == 0 or c == 3 for c in (conv.in_channels, conv.out_channels))


class _ConvFn(torch.autograd.Function):
    """y = conv(x, weight, bias), k in {1, 3}, padding k // 2, stride in {1, 2}.  `cache`: a _Packed (weights of an
    nn.Conv2d: packings reused until the weight changes) or None (weights recomputed every step, e.g. GDN's gamma)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, cache):
        eng = train_engine()
        n, cin, h, w = x.shape
        cout, _, k, _ = weight.shape
        xm = _nhwc(x)
        hi, lo = _split(eng, xm, n, h, w, cin)
        pc = cache.fwd if cache is not None else None
        if pc is None:
            pc = pack_conv(weight, bias, stride, _lib.STORE_NHWC, weight.device,
                           exp=None if cache is None else cache.exp_fwd)
            if cache is not None:
                cache.fwd, cache.exp_fwd = pc, _exp_of(pc)
        out = eng.conv(pc, (hi, lo), Act(n, h, w, pc.cin), {"f32"})
        y = out.f32                                         # [n, ho, wo, cout_pad8]
        ctx.cache, ctx.stride = cache, stride
        ctx.shape = (n, cin, h, w)
        ctx.save_for_backward(hi, weight)
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)[:, :cout] if y.shape[-1] != cout else y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        eng = train_engine()
        hi, weight = ctx.saved_tensors
        cache, stride = ctx.cache, ctx.stride
        n, cin, h, w = ctx.shape
        cout, _, k, _ = weight.shape
        ho, wo = h // stride, w // stride
        gm = _nhwc(g)
        s, inv_s = _grad_scale(gm)
        gp, gl = _split(eng, gm, n, ho, wo, cout, dev_scale=s)      # [n, ho, wo, cout_pad8] fp16 of g * S
        cop = gp.shape[-1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            pd = cache.dgrad if cache is not None else None
            if pd is None:
                wd = _dgrad_weight(weight, stride)
                pd = pack_conv(wd, None, 1, _lib.STORE_SHUFFLE_NHWC if stride == 2 else _lib.STORE_NHWC, weight.device,
                               exp=None if cache is None else cache.exp_dgrad)
                if cache is not None:
                    cache.dgrad, cache.exp_dgrad = pd, _exp_of(pd)
            out = eng.conv(pd, (gp, gl), Act(n, ho, wo, cop), {"f32"}, dev_scale=inv_s)
            d = out.f32                                     # [n, h, w, cin_pad8]
            dx = d.permute(0, 3, 1, 2)
            if d.shape[-1] != cin:
                dx = dx[:, :cin]
        if ctx.needs_input_grad[1]:
            cip = hi.shape[-1]
            dwp = torch.empty((cop, cip, k, k), dtype=torch.float32, device=g.device)
            p = _lib.WgradParams()
            p.x_hi, p.n, p.hin, p.win, p.cin = _ptr(hi), n, h, w, cip
            p.dy_hi, p.cout, p.ksize, p.stride = _ptr(gp), cop, k, stride
            p.dw, p.scale, p.dev_scale, p.accumulate = _ptr(dwp), 1.0, _ptr(inv_s), 0
            need = int(eng.lib.mcq_conv_wgrad_workspace_bytes(ctypes.byref(p)))
            if need < 0:
                raise RuntimeError("mcquic_b200: convolution shape not supported by mcq_conv_wgrad")
            ws = torch.empty(need + 256, dtype=torch.uint8, device=g.device)
            off = (-ws.data_ptr()) % 256
            p.workspace, p.workspace_bytes = ctypes.c_void_p(ws.data_ptr() + off), need
            _lib.check(eng.lib.mcq_conv_wgrad(ctypes.byref(p), eng._stream()), "mcq_conv_wgrad")
            dw = dwp[:cout, :cin] if (cop, cip) != (cout, cin) else dwp
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = g.sum((0, 2, 3))
        return dx, dw, db, None, None


def conv2d_weights(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1,
                   owner: Optional[nn.Module] = None) -> torch.Tensor:
    """convolution with explicitly given weights [cout, cin, k, k] (k in {1, 3}, padding k // 2) on the tcgen05 kernels;
    owner: the module the weights are derived from (its operand exponents are cached, see _Packed)"""
    return _ConvFn.apply(x, weight, bias, stride, None if owner is None else _packs_for(owner, weight, bias))


def conv2d(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d `conv` applied to x [n, c, h, w] (fp32, CUDA) with forward, dgrad and wgrad on the tcgen05 kernels"""
    if not x.is_cuda:
        raise RuntimeError("mcquic_b200 runs on CUDA tensors only (no CPU fallback)")
    if conv_supported(conv):
        return _ConvFn.apply(x, conv.weight, conv.bias, conv.stride[0], _packs_for(conv, conv.weight, conv.bias))
    return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding)


# ------------------------------------------------------------------------------------------------------------------
# blocks (mcquic/nn/blocks.py, gdn.py, base.py) as differentiable functions of the parameter containers in mcquic_b200.nn
class _LowerBoundFn(torch.autograd.Function):
    """mcquic/nn/base.py:17-29: max(x, bound) whose gradient passes where x >= bound or the gradient pushes x up"""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)).type(g.dtype) * g, None


def _reparam(rep, p: torch.Tensor) -> torch.Tensor:
    """NonNegativeParametrizer.forward (base.py:81-84)"""
    return _LowerBoundFn.apply(p, rep.lowerBound.bound) ** 2 - rep.eps


def gdn(mod, x: torch.Tensor) -> torch.Tensor:
    """GenDivNorm / InvGenDivNorm.forward (gdn.py:67-91): the 1x1 convolution over x^2 runs on the tcgen05 kernels"""
    beta = _reparam(mod.beta_reparam, mod.beta)
    gamma = _reparam(mod.gamma_reparam, mod.gamma)[..., None, None]
    # the operand plane holds x^2 * 2^-6 like the inference path's MCQ_ACT_SQUARE planes (|x| up to 2047 stays in fp16 range)
    std = conv2d_weights(x ** 2 * _lib.SQUARE_SCALE, gamma, None, owner=mod) * (1.0 / _lib.SQUARE_SCALE) \
        + beta.reshape(1, -1, 1, 1)
    return x * torch.sqrt(std) if mod.inverse else x * torch.rsqrt(std)


def run_block(mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
    from .nn.blocks import AttentionBlock, _ResidualBase
    from .nn.gdn import GenDivNorm
    if isinstance(mod, nn.Conv2d):
        return conv2d(mod, x)
    if isinstance(mod, _ResidualBase):                       # blocks.py:62-78
        b = mod._branch
        out = run_block(b[1], F.silu(x))
        if isinstance(b[2], GenDivNorm):
            out = gdn(b[2], out)
        elif isinstance(b[2], nn.GroupNorm):
            out = F.group_norm(out, b[2].num_groups, b[2].weight, b[2].bias, b[2].eps)
        else:
            out = F.silu(out)
        out = run_block(b[3], out)
        return out + (x if mod._skip is None else run_block(mod._skip, x))
    if isinstance(mod, AttentionBlock):                      # blocks.py:281-288
        a = run_block(mod._mainBranch, x)
        b = run_block(mod._sideBranch, x)
        return a * torch.sigmoid(b) + x
    if isinstance(mod, nn.Sequential):
        for sub in mod:
            x = run_block(sub, x)
        return x
    if isinstance(mod, nn.PixelShuffle):
        return F.pixel_shuffle(x, mod.upscale_factor)
    if isinstance(mod, nn.Identity):
        return x
    raise NotImplementedError(f"mcquic_b200: no training path for {type(mod).__name__}")


# ------------------------------------------------------------------------------------------------------------------
# soft quantizer (mcquic/modules/quantizer.py:181-274)
class _LogitsFn(torch.autograd.Function):
    """logit[n, m, h, w, k] = -(|x|^2 + |c_k|^2 - 2 x.c_k) / sqrt(k) * t[m]  (quantizer.py:153-183,204).  Forward: the VQ
    launch of the inference path (mcq_vq_assign / the fused tcgen05 kernel).  Backward (this round): the closed form below
    evaluated with torch.bmm -- the VQ is 0.6 % of the step's FLOPs."""

    @staticmethod
    def forward(ctx, x, codebook, t):
        eng = train_engine()
        n, c, h, w = x.shape
        m, k, d = codebook.shape
        xm = _nhwc(x)
        cb = codebook.detach().float().contiguous()
        c2 = (cb ** 2).sum(-1).contiguous()
        scale = t.detach().float().reshape(-1).contiguous()
        _, logit = eng.vq_assign(xm, cb, c2, n, h, w, logits=True, logit_scale=scale)
        ctx.save_for_backward(x, codebook, t, logit)
        return logit

    @staticmethod
    def backward(ctx, g):
        x, codebook, t, logit = ctx.saved_tensors
        n, c, h, w = x.shape
        m, k, d = codebook.shape
        coef = (-t.reshape(m) / (k ** 0.5)).float()                       # logit = coef * dist
        G = g.reshape(n, m, h * w, k).float()
        X = x.reshape(n, m, d, h * w).permute(0, 1, 3, 2).float()            # [n, m, hw, d]
        cb = codebook.float()
        dx = dc = dt = None
        if ctx.needs_input_grad[0]:
            gs = G.sum(-1, keepdim=True)                                      # [n, m, hw, 1]
            gc = torch.matmul(G, cb[None])                                    # [n, m, hw, d]
            dxp = 2.0 * coef.reshape(1, m, 1, 1) * (X * gs - gc)
            dx = dxp.permute(0, 1, 3, 2).reshape(n, c, h, w)
        if ctx.needs_input_grad[1]:
            gk = G.sum((0, 2))                                                # [m, k]
            gx = torch.einsum("nmpk,nmpd->mkd", G, X)
            dc = 2.0 * coef.reshape(m, 1, 1) * (cb * gk[..., None] - gx)
        if ctx.needs_input_grad[2]:
            dt = ((g * logit).sum((0, 2, 3, 4)) / t.reshape(m)).reshape(t.shape)
        return dx, dc, dt


def _gumbel_softmax_hard(logits: torch.Tensor) -> torch.Tensor:
    """mcquic/nn/base.py:118-133 with temperature 1, hard = True (straight-through)"""
    eps = torch.finfo(logits.dtype).eps
    uniforms = torch.rand_like(logits).clamp_(eps, 1 - eps)
    gumbels = -((-(uniforms.log())).log())
    y_soft = (logits + gumbels).softmax(-1)
    index = y_soft.max(-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


def quantize_soft(q, x: torch.Tensor):
    """_multiCodebookQuantization.forward (quantizer.py:202-239): (sample, code, oneHot, logit)"""
    t = _LowerBoundFn.apply(q._temperature, q._bound.bound)                   # quantizer.py:204
    logit = _LogitsFn.apply(x, q._codebook, t)
    if isinstance(getattr(q, "_freqEMA", None), torch.Tensor):
        logit = q._randomDrop(logit)        # quantizer.py:194-200 (out of place: the un-masked logits feed dL/dtemperature)
    sample = _gumbel_softmax_hard(logit)
    code = logit.argmax(-1, keepdim=True)
    one_hot = torch.zeros_like(logit).scatter_(-1, code, 1).contiguous()
    return sample, code[..., 0].contiguous(), one_hot, logit


def dequantize_soft(dq, sample: torch.Tensor) -> torch.Tensor:
    """_multiCodebookDeQuantization.forward (quantizer.py:262-274)"""
    n, m, h, w, k = sample.shape
    left = sample.reshape(n * m, h * w, k)
    right = dq._codebook.expand(n, m, k, dq._d).reshape(n * m, k, dq._d)
    return torch.bmm(left, right).reshape(n, m, h, w, dq._d).permute(0, 1, 4, 2, 3).reshape(n, -1, h, w)


def quantizer_forward(quantizer, y: torch.Tensor):
    """UMGMQuantizer.forward (quantizer.py:443-467) / ResidualBackwardQuantizer.forward (:727-765) -> (yHat, codes, logits)"""
    from .modules.quantizer import ResidualBackwardQuantizer
    codes, one_hots, logits, quantizeds = [], [], [], []
    if isinstance(quantizer, ResidualBackwardQuantizer):
        latents, x = [], y
        for enc in quantizer._encoders:
            x = run_block(enc, x)
            latents.append(x)
        current = torch.zeros_like(latents[-1])
        for lv in reversed(range(len(latents))):
            sample, code, one_hot, logit = quantize_soft(quantizer._quantizers[lv], latents[lv] - current)
            quantized = dequantize_soft(quantizer._dequantizers[lv], sample)
            quantizeds.append(quantized)
            codes.append(code)
            one_hots.append(one_hot)
            logits.append(logit)
            current = run_block(quantizer._backwards[lv], quantized)
        former = torch.zeros_like(quantizeds[0])
        for lv, quantized in zip(reversed(range(len(latents))), quantizeds):
            former = run_block(quantizer._decoders[lv], former + quantized)
    else:
        x = y
        for enc in quantizer._encoders:                                       # quantizer.py:320-328
            z = run_block(enc._latentStageEncoder, x)
            sample, code, one_hot, logit = quantize_soft(enc._quantizer, run_block(enc._quantizationHead, z))
            if enc._latentHead is not None:
                x = run_block(enc._latentHead, z) - dequantize_soft(enc._dequantizer, sample)
            quantizeds.append(sample)
            codes.append(code)
            one_hots.append(one_hot)
            logits.append(logit)
        former = None
        for dec, sample in zip(quantizer._decoders[::-1], quantizeds[::-1]):  # quantizer.py:359-365
            q = run_block(dec._dequantizationHead, dequantize_soft(dec._dequantizer, sample))
            if dec._sideHead is not None:
                q = q + run_block(dec._sideHead, former)
            former = run_block(dec._restoreHead, q)
    quantizer._entropyCoder([o.detach() for o in one_hots])                   # frequency EMA (+ all-reduce), no gradient
    return former, codes, logits


def compressor_forward(model, x: torch.Tensor):
    """BaseCompressor.forward (compressor.py:35-43) with an autograd graph: (xHat, yHat, codes, logits)"""
    if not x.is_cuda:
        raise RuntimeError("mcquic_b200 runs on CUDA tensors only (no CPU fallback)")
    y = run_block(model._encoder, x.float())
    yHat, codes, logits = quantizer_forward(model._quantizer, y)
    xHat = run_block(model._decoder, yHat)
    return xHat, yHat, codes, logits
