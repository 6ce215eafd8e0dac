"""Shared helper for the conv parity tests: build one mcq_conv2d problem from seeded data, run it through the
engine (CUDA) and compute the expected result in fp64 with torch from the *same* split-fp16 operands."""
import torch
import torch.nn.functional as F

from mcquic_b200 import _lib
from mcquic_b200.engine import Act, Engine, pack_conv

LO = 2048.0


def make_planes(x_nhwc: torch.Tensor, passes: int):
    hi = x_nhwc.clamp(-65504, 65504).half()
    lo = ((x_nhwc - hi.float()) * LO).half() if passes == 3 else None
    return hi.contiguous(), (lo.contiguous() if lo is not None else None)


def planes_value(pl, passes):
    v = pl[0].double()
    if passes == 3 and pl[1] is not None:
        v = v + pl[1].double() / LO
    return v


def run_case(eng: Engine, *, n, h, w, cin, cout, ksize=3, stride=1, passes=3, store=_lib.STORE_NHWC,
             mode=_lib.EPI_LINEAR, use_res1=False, res1_scale=1.0, use_res2=False, seed=0, want=("f32",),
             device="cuda", amp=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(n, h, w, cin, generator=g) * amp).to(device)
    fan = cin * ksize * ksize
    wt = (torch.rand(cout, cin, ksize, ksize, generator=g) * 2 - 1).to(device) / fan ** 0.5
    bias = (torch.rand(cout, generator=g) * 2 - 1).to(device) / fan ** 0.5
    if mode in (_lib.EPI_GDN, _lib.EPI_IGDN):
        wt = wt.abs() + 0.01
        bias = bias.abs() + 0.5
        x = x * x  # operand of a GDN conv is a square
    pc = pack_conv(wt, bias, stride, store, device)
    eng.passes = passes
    a = make_planes(x, passes)
    act = Act(n, h, w, cin)
    ho, wo = h // stride, w // stride
    co = cout
    if store != _lib.STORE_NHWC:
        ho, wo, co = 2 * ho, 2 * wo, cout // 4
    oshape = (n, ho, wo, co)
    res1 = torch.randn(oshape, generator=g).to(device) if (use_res1 or mode == _lib.EPI_GATE) else None
    res2 = torch.randn(oshape, generator=g).to(device) if use_res2 else None
    aux = torch.randn(oshape, generator=g).to(device) if mode != _lib.EPI_LINEAR else None
    out = eng.conv(pc, a, act, set(want), mode=mode, res1=res1, res1_scale=res1_scale, res2=res2, aux=aux)
    eng.flush()   # small maps are queued for a layer-chain launch
    torch.cuda.synchronize() if device == "cuda" else None

    # ---- expected, fp64, from the same quantised operands
    av = planes_value(a, passes).permute(0, 3, 1, 2)
    wv = planes_value((pc.w_hi, pc.w_lo), passes)[:cout].reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    acc = F.conv2d(av, wv, None, stride=stride, padding=ksize // 2) * pc.w_scale
    v = acc + pc.bias.double()[None, :, None, None]
    if store == _lib.STORE_SHUFFLE_NCHW:
        exp = F.pixel_shuffle(v, 2)
        return out, {"f32": exp}, pc
    y = v.permute(0, 2, 3, 1)
    if store == _lib.STORE_SHUFFLE_NHWC:
        y = y.reshape(n, h // stride, w // stride, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(oshape)
    if mode == _lib.EPI_LINEAR:
        if res1 is not None:
            y = y + res1_scale * res1.double()
        if res2 is not None:
            y = y + res2.double()
    elif mode == _lib.EPI_GATE:
        y = aux.double() * torch.sigmoid(y) + res1.double()
    elif mode == _lib.EPI_GDN:
        y = aux.double() / torch.sqrt(y)
    else:
        y = aux.double() * torch.sqrt(y)
    exp = {"f32": y, "raw": y, "silu": F.silu(y), "sq": y * y * _lib.SQUARE_SCALE}
    return out, exp, pc


def compare(out: Act, exp, want, passes):
    """max abs error per representation, relative to the max magnitude of the expectation"""
    res = {}
    for name in want:
        e = exp[name]
        if name == "f32":
            got = out.f32.double()
        else:
            got = planes_value(getattr(out, name), passes)
        scale = float(e.abs().max()) + 1e-30
        res[name] = float((got - e).abs().max()) / scale
    return res
