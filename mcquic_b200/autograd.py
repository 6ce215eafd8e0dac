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
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .engine import Act, Engine, pack_conv, _ptr

_ENGINE: Optional[Engine] = None


def train_engine() -> Engine:
    """the engine the training-step convolutions launch through: one fp16 pass, single stream (autograd orders the work)"""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine()
        _ENGINE.multistream = False
        _ENGINE.chain = False
    _ENGINE.passes = 1
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
           dev_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 tensor whose memory is [n, h, w, c] -> fp16 plane [n, h, w, c_pad] of x * (*dev_scale)"""
    if c % pad_to != 0:      # RGB input / 3-channel output gradient: zero channels up to the vector width of the kernels
        cp = (c + pad_to - 1) // pad_to * pad_to
        padded = x_nhwc_mem.new_zeros((n, cp, h, w)).contiguous(memory_format=torch.channels_last)
        padded[:, :c] = x_nhwc_mem
        x_nhwc_mem, c = padded, cp
    hi = torch.empty((n, h, w, c), dtype=torch.float16, device=x_nhwc_mem.device)
    _lib.check(eng.lib.mcq_split_planes(_ptr(x_nhwc_mem), x_nhwc_mem.numel(), _lib.ACT_NONE, _ptr(hi), None,
                                        _ptr(dev_scale), eng._stream()), "mcq_split_planes")
    return hi


def _grad_scale(g: torch.Tensor):
    """(S, 1/S) as device scalars: S = the power of two that brings max |g| to [256, 512) -- fp16 planes then keep ~2^-23
    of the largest gradient before flushing.  A handful of tiny launches, no host synchronisation."""
    amax = g.detach().abs().amax().float()
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
    """per-convolution cache of the forward / dgrad operand packings, keyed on the weight's version counter"""

    def __init__(self):
        self.key = None
        self.fwd = None
        self.dgrad = None


_PACKS = {}


def _packs_for(conv: nn.Conv2d) -> _Packed:
    pk = _PACKS.get(id(conv))
    key = (_version(conv.weight), conv.weight.data_ptr(), None if conv.bias is None else _version(conv.bias))
    if pk is None or pk.key != key:
        pk = _PACKS[id(conv)] = _Packed()
        pk.key = key
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
    @staticmethod
    def forward(ctx, x, weight, bias, conv):
        eng = train_engine()
        n, cin, h, w = x.shape
        stride, k = conv.stride[0], conv.kernel_size[0]
        cout = conv.out_channels
        xm = _nhwc(x)
        hi = _split(eng, xm, n, h, w, cin)
        pk = _packs_for(conv)
        if pk.fwd is None:
            pk.fwd = pack_conv(weight, bias, stride, _lib.STORE_NHWC, weight.device)
        pc = pk.fwd
        out = eng.conv(pc, (hi, None), Act(n, h, w, pc.cin), {"f32"})
        y = out.f32                                         # [n, ho, wo, cout_pad8]
        ctx.conv = conv
        ctx.shape = (n, cin, h, w)
        ctx.save_for_backward(hi, weight)
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)[:, :cout] if y.shape[-1] != cout else y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        eng = train_engine()
        hi, weight = ctx.saved_tensors
        conv = ctx.conv
        n, cin, h, w = ctx.shape
        stride, k = conv.stride[0], conv.kernel_size[0]
        cout = conv.out_channels
        ho, wo = h // stride, w // stride
        gm = _nhwc(g)
        s, inv_s = _grad_scale(gm)
        gp = _split(eng, gm, n, ho, wo, cout, dev_scale=s)          # [n, ho, wo, cout_pad8] fp16 of g * S
        cop = gp.shape[-1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            pk = _packs_for(conv)
            if pk.dgrad is None:
                wd = _dgrad_weight(weight, stride)
                pk.dgrad = pack_conv(wd, None, 1, _lib.STORE_SHUFFLE_NHWC if stride == 2 else _lib.STORE_NHWC, weight.device)
            out = eng.conv(pk.dgrad, (gp, None), Act(n, ho, wo, cop), {"f32"}, dev_scale=inv_s)
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
        return dx, dw, db, None


def conv2d(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d `conv` applied to x [n, c, h, w] (fp32, CUDA) with forward, dgrad and wgrad on the tcgen05 kernels"""
    if not x.is_cuda:
        raise RuntimeError("mcquic_b200 runs on CUDA tensors only (no CPU fallback)")
    if conv_supported(conv):
        return _ConvFn.apply(x, conv.weight, conv.bias, conv)
    return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding)
