"""CPU model of the C ABI in include/mcquic_b200.h  --  TEST INFRASTRUCTURE ONLY.

Implements every entry point on host memory with numpy/torch so that the *host-side* logic of
mcquic_b200 (module walk, fusion map, weight repacking / PixelShuffle row permutation, stride-2 tap
tables, split-fp16 plane format) can be checked against the oracle on a box without a GPU.  It is written
from the header's contract, not from the CUDA sources, so it doubles as a second reading of the spec.
Never imported by the product.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from mcquic_b200 import _lib

LO = 2048.0


def _arr(ptr, shape, dtype):
    if not ptr:
        return None
    count = int(np.prod(shape))
    ctype = {np.float32: ctypes.c_float, np.float16: ctypes.c_uint16, np.int64: ctypes.c_int64,
             np.int32: ctypes.c_int32}[dtype]
    buf = (ctype * count).from_address(ptr)
    a = np.ctypeslib.as_array(buf)
    if dtype is np.float16:
        a = a.view(np.float16)
    return a.reshape(shape)


def _act(y, act):
    if act == _lib.ACT_SILU:
        return F.silu(y)
    if act == _lib.ACT_SQUARE:
        return y * y * _lib.SQUARE_SCALE
    return y


def _store_planes(hi_ptr, lo_ptr, shape, y, act):
    if not hi_ptr:
        return
    t = _act(y, act).float()
    hi = t.clamp(-65504, 65504).half()
    _arr(hi_ptr, shape, np.float16)[...] = hi.numpy()
    if lo_ptr:
        lo = ((t - hi.float()) * LO).clamp(-65504, 65504).half()
        _arr(lo_ptr, shape, np.float16)[...] = lo.numpy()


class EmulatedLib:
    """duck-types the ctypes CDLL object `mcquic_b200._lib.load()` returns"""

    def __init__(self):
        self.launches = 0

    # ------------------------------------------------------------------
    def mcq_conv2d(self, pref, stream):
        p = pref._obj
        n, hin, win, cin = p.n, p.hin, p.win, p.cin
        k, s = p.ksize, p.stride
        K = k * k * cin
        a = torch.from_numpy(_arr(p.a_hi, (n, hin, win, cin), np.float16).astype(np.float32))
        w = torch.from_numpy(_arr(p.w_hi, (p.cout_pad, K), np.float16).astype(np.float32))
        if p.passes == 3:
            a = a.double() + torch.from_numpy(_arr(p.a_lo, (n, hin, win, cin), np.float16).astype(np.float64)) / LO
            w = w.double() + torch.from_numpy(_arr(p.w_lo, (p.cout_pad, K), np.float16).astype(np.float64)) / LO
        else:
            a, w = a.double(), w.double()
        w4 = w[:p.cout].reshape(p.cout, k, k, cin).permute(0, 3, 1, 2)
        acc = F.conv2d(a.permute(0, 3, 1, 2), w4, None, stride=s, padding=k // 2)  # [n, cout, ho, wo]
        bias = torch.from_numpy(_arr(p.bias, (p.cout,), np.float32).copy())
        v = (acc * p.w_scale).float() + bias[None, :, None, None]
        ho, wo = hin // s, win // s
        if p.store == _lib.STORE_SHUFFLE_NCHW:
            pix = F.pixel_shuffle(v, 2)
            if p.out_u8:
                u8 = (((pix - (-1.0)) / 2.0) * 255.999).clamp(0.0, 255.0).byte()      # DeTransform, vision.py:135-146
                buf = (ctypes.c_uint8 * u8.numel()).from_address(p.out_u8)
                np.ctypeslib.as_array(buf).reshape(tuple(u8.shape))[...] = u8.numpy()
            else:
                _arr(p.out_f32, (n, p.cout // 4, 2 * ho, 2 * wo), np.float32)[...] = pix.numpy()
            self.launches += 1
            return 0
        y = v.permute(0, 2, 3, 1)  # NHWC, GEMM-column order
        if p.store == _lib.STORE_SHUFFLE_NHWC:
            cq = p.cout // 4
            # column (2i+j)*cq + c -> pixel (2y+i, 2x+j), channel c
            y = y.reshape(n, ho, wo, 2, 2, cq).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * ho, 2 * wo, cq)
        shape = tuple(y.shape)
        g = lambda ptr: None if not ptr else torch.from_numpy(_arr(ptr, shape, np.float32).copy())
        res1, res2, aux = g(p.res1), g(p.res2), g(p.aux)
        if p.mode == _lib.EPI_LINEAR:
            if res1 is not None:
                y = y + p.res1_scale * res1
            if res2 is not None:
                y = y + res2
        elif p.mode == _lib.EPI_GATE:
            y = aux * torch.sigmoid(y) + res1
        elif p.mode == _lib.EPI_GDN:
            y = aux * (1.0 / torch.sqrt(y))
        else:
            y = aux * torch.sqrt(y)
        y = y.contiguous()
        if p.out_f32:
            _arr(p.out_f32, shape, np.float32)[...] = y.numpy()
        if p.gn_partials:
            # header contract: (sum, sum^2) per (row block, unit channels); this model uses ONE row block per image
            unit = min(p.cout // p.gn_groups, 16)
            yu = y.double().reshape(n, -1, p.cout // unit, unit)
            part = torch.stack([yu.sum((1, 3)), (yu * yu).sum((1, 3))], -1).float()     # [n, units, 2]
            _arr(p.gn_partials, (n, 1, p.cout // unit, 2), np.float32)[...] = part[:, None].numpy()
        _store_planes(p.out0_hi, p.out0_lo, shape, y, p.out0_act)
        _store_planes(p.out1_hi, p.out1_lo, shape, y, p.out1_act)
        self.launches += 1
        return 0

    def mcq_conv_chain(self, params, count, stream):
        """header contract: same as calling mcq_conv2d on params[0..count) in order"""
        self.chains = getattr(self, "chains", 0) + 1
        for i in range(count):
            rc = self.mcq_conv2d(ctypes.byref(params[i]), stream)
            if rc:
                return rc
        return 0

    def mcq_conv_chain_max_layers(self):
        return 28

    def mcq_stem_conv(self, x, x_is_u8, n, h, w, top, left, hp, wp, wgt, bias, cout, out_f32, out_hi, out_lo, act, stream):
        if x_is_u8:
            buf = (ctypes.c_uint8 * (n * 3 * h * w)).from_address(x.value)
            xi = torch.from_numpy(np.ctypeslib.as_array(buf).reshape(n, 3, h, w).copy())
            xi = (xi.float() / 255.0 - 0.5) * 2          # demo.py:110-118
        else:
            xi = torch.from_numpy(_arr(x.value, (n, 3, h, w), np.float32).copy())
        if hp != h or wp != w:
            xi = F.pad(xi, (left, wp - w - left, top, hp - h - top), "reflect")
        wt = torch.from_numpy(_arr(wgt.value, (cout, 3, 3, 3), np.float32).copy())
        b = torch.from_numpy(_arr(bias.value, (cout,), np.float32).copy())
        y = F.conv2d(xi, wt, b, stride=2, padding=1).permute(0, 2, 3, 1).contiguous()
        shape = tuple(y.shape)
        if out_f32 is not None and out_f32.value:
            _arr(out_f32.value, shape, np.float32)[...] = y.numpy()
        _store_planes(out_hi.value if out_hi is not None else 0, out_lo.value if out_lo is not None else 0, shape, y, act)
        self.launches += 1
        return 0

    def mcq_vq_assign(self, x, codebook, c2, codes, logits, logit_scale, hist, n, h, w, m, k, d, stream):
        xi = torch.from_numpy(_arr(x.value, (n * h * w, m, d), np.float32).copy())
        cb = torch.from_numpy(_arr(codebook.value, (m, k, d), np.float32).copy())
        cc = torch.from_numpy(_arr(c2.value, (m, k), np.float32).copy())
        x2 = (xi ** 2).sum(-1)                                    # [P, m]
        inter = torch.einsum("pmd,mkd->pmk", xi, cb)
        dist = (x2[..., None] + cc[None]) - 2 * inter             # [P, m, k]
        code = dist.argmin(-1).reshape(n, h * w, m).permute(0, 2, 1).reshape(n, m, h, w)
        _arr(codes.value, (n, m, h, w), np.int64)[...] = code.numpy()
        if logits is not None and logits.value:
            sc = torch.ones(m) if logit_scale is None or not logit_scale.value else torch.from_numpy(
                _arr(logit_scale.value, (m,), np.float32).copy())
            lg = (-dist / (k ** 0.5)) * sc[None, :, None]
            _arr(logits.value, (n, m, h, w, k), np.float32)[...] = lg.reshape(n, h, w, m, k).permute(0, 3, 1, 2, 4).numpy()
        if hist is not None and hist.value:
            hv = _arr(hist.value, (m, k), np.int32)
            for j in range(m):
                hv[j] += np.bincount(code[:, j].flatten().numpy(), minlength=k).astype(np.int32)
        self.launches += 1
        return 0

    def mcq_vq_dequant(self, codes, codebook, n, h, w, m, k, d, out_f32, o0h, o0l, a0, o1h, o1l, a1, status, stream):
        code = torch.from_numpy(_arr(codes.value, (n, m, h, w), np.int64).copy())
        cb = torch.from_numpy(_arr(codebook.value, (m, k, d), np.float32).copy())
        if ((code < 0) | (code >= k)).any():
            if status is not None and status.value:
                _arr(status.value, (1,), np.int32)[0] = -4
            code = code.clamp(0, k - 1) * ((code >= 0) & (code < k))
        ix = torch.arange(m)[None, None, None, :].expand(n, h, w, m)
        y = cb[ix, code.permute(0, 2, 3, 1)].reshape(n, h, w, m * d).contiguous()
        shape = tuple(y.shape)
        v = lambda q: q.value if q is not None and q.value else 0
        if v(out_f32):
            _arr(v(out_f32), shape, np.float32)[...] = y.numpy()
        _store_planes(v(o0h), v(o0l), shape, y, a0)
        _store_planes(v(o1h), v(o1l), shape, y, a1)
        self.launches += 1
        return 0

    def mcq_code_histogram(self, codes, n, m, hw, k, hist, stream):
        code = _arr(codes.value, (n, m, hw), np.int64)
        hv = _arr(hist.value, (m, k), np.int32)
        for j in range(m):
            c = code[:, j].reshape(-1)
            c = c[(c >= 0) & (c < k)]
            hv[j] += np.bincount(c, minlength=k).astype(np.int32)
        self.launches += 1
        return 0

    def mcq_groupnorm(self, x, n, h, w, c, groups, gamma, beta, eps, out_f32, out_hi, out_lo, act, stream):
        """header contract: nn.GroupNorm(groups, c) on fp32 NHWC, outputs fp32 and/or planes of act(y)"""
        xi = torch.from_numpy(_arr(x.value, (n, h, w, c), np.float32).copy())
        g = torch.from_numpy(_arr(gamma.value, (c,), np.float32).copy())
        b = torch.from_numpy(_arr(beta.value, (c,), np.float32).copy())
        y = F.group_norm(xi.permute(0, 3, 1, 2), groups, g, b, eps).permute(0, 2, 3, 1).contiguous()
        v = lambda q: q.value if q is not None and q.value else 0
        if v(out_f32):
            _arr(v(out_f32), (n, h, w, c), np.float32)[...] = y.numpy()
        _store_planes(v(out_hi), v(out_lo), (n, h, w, c), y, act)
        self.launches += 1
        return 0

    def mcq_conv_gn_layout(self, pref, rb_ref, unit_ref):
        p = pref._obj
        cg = p.cout // p.gn_groups
        if p.ksize != 3 or p.stride != 1 or p.store != _lib.STORE_NHWC or not p.out_f32 or cg % 4 or (cg > 16 and cg % 16) \
                or (cg < 16 and 16 % cg) or p.cin % 64:
            return _lib.ERR_UNSUPPORTED
        rb_ref._obj.value, unit_ref._obj.value = 1, min(cg, 16)
        return 0

    def mcq_groupnorm_apply(self, x, partials, rb, unit, n, h, w, c, groups, gamma, beta, eps, stats, out_f32, out_hi,
                            out_lo, act, stream):
        xi = torch.from_numpy(_arr(x.value, (n, h, w, c), np.float32).copy())
        part = torch.from_numpy(_arr(partials.value, (n, rb, c // unit, 2), np.float32).copy()).double().sum(1)
        cg = c // groups
        sums = part.reshape(n, groups, cg // unit, 2).sum(2)                # [n, groups, 2]
        cnt = h * w * cg
        mean = sums[..., 0] / cnt
        var = (sums[..., 1] / cnt - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        _arr(stats.value, (n, groups, 2), np.float32)[...] = torch.stack([mean, rstd], -1).float().numpy()
        g = torch.from_numpy(_arr(gamma.value, (c,), np.float32).copy())
        b = torch.from_numpy(_arr(beta.value, (c,), np.float32).copy())
        m_c = mean.float().repeat_interleave(cg, 1)[:, None, None, :]
        r_c = rstd.float().repeat_interleave(cg, 1)[:, None, None, :]
        y = ((xi - m_c) * r_c * g + b).contiguous()
        v = lambda q: q.value if q is not None and q.value else 0
        if v(out_f32):
            _arr(v(out_f32), (n, h, w, c), np.float32)[...] = y.numpy()
        _store_planes(v(out_hi), v(out_lo), (n, h, w, c), y, act)
        self.launches += 1
        return 0

    def mcq_add_scaled(self, x, y, alpha, count, out_f32, out_hi, out_lo, act, stream):
        a = torch.from_numpy(_arr(x.value, (count,), np.float32).copy())
        b = torch.from_numpy(_arr(y.value, (count,), np.float32).copy())
        v = (a.double() + float(alpha) * b.double()).float()      # one rounding, like the kernel's FFMA
        q = lambda t: t.value if t is not None and t.value else 0
        if q(out_f32):
            _arr(q(out_f32), (count,), np.float32)[...] = v.numpy()
        _store_planes(q(out_hi), q(out_lo), (count,), v, act)
        self.launches += 1
        return 0

    def mcq_split_planes(self, x, count, act, out_hi, out_lo, dev_scale, stream):
        y = torch.from_numpy(_arr(x.value, (count,), np.float32).copy())
        _store_planes(out_hi.value, out_lo.value if out_lo is not None and out_lo.value else 0, (count,), y, act)
        self.launches += 1
        return 0

    def mcq_nchw_to_nhwc(self, x, n, c, h, w, out_f32, o0h, o0l, a0, o1h, o1l, a1, stream):
        y = torch.from_numpy(_arr(x.value, (n, c, h, w), np.float32).copy()).permute(0, 2, 3, 1).contiguous()
        shape = tuple(y.shape)
        v = lambda q: q.value if q is not None and q.value else 0
        if v(out_f32):
            _arr(v(out_f32), shape, np.float32)[...] = y.numpy()
        _store_planes(v(o0h), v(o0l), shape, y, a0)
        _store_planes(v(o1h), v(o1l), shape, y, a1)
        self.launches += 1
        return 0

    def mcq_nhwc_to_nchw(self, x, n, c, h, w, out, stream):
        y = torch.from_numpy(_arr(x.value, (n, h, w, c), np.float32).copy()).permute(0, 3, 1, 2).contiguous()
        _arr(out.value, (n, c, h, w), np.float32)[...] = y.numpy()
        self.launches += 1
        return 0

    def mcq_error_string(self, code):
        return f"emulated error {code}".encode()

    def mcq_kernel_launch_count(self):
        return self.launches
