"""Multi-codebook quantizer with the reference's classes, method names and state_dict layout
(mcquic/modules/quantizer.py:99-274 _multiCodebookQuantization / _multiCodebookDeQuantization,
:277-365 _quantizerEncoder / _quantizerDecoder, :368-467 UMGMQuantizer) on the CUDA engine.

State-dict keys kept: `_encoders.{l}._quantizer.{_codebook,_temperature}`, `_encoders.{l}._dequantizer._codebook`,
`_decoders.{l}._dequantizer._codebook` (the three codebook keys alias one Parameter, as upstream),
`_encoders.{l}.{_latentStageEncoder,_quantizationHead,_latentHead}.*`,
`_decoders.{l}.{_dequantizationHead,_sideHead,_restoreHead}.*`, `_entropyCoder._freqEMA.{l}`.
"""
import math
from typing import Callable, Dict, List, Optional, Union

import torch
from torch import nn

from .. import _lib
from ..engine import Act, Engine, default_engine
from ..nn.gdn import LowerBound

_EPS = 1e-6  # mcquic/consts.py:25


class CodeFrequency(nn.Module):
    """The reference's `EntropyCoder` (mcquic/modules/entropyCoder.py:15-154): per-level [m, k] frequency EMA (initially
    uniform; checkpoint keys `_entropyCoder._freqEMA.{l}`), quantized CDF tables and per-image rANS streams.  The
    coder itself is the host C++ library behind `mcquic_b200.entropy` (bit-identical streams to `mcquic.rans`)."""

    def __init__(self, m: int, k: List[int], ema: float = 0.9):
        super().__init__()
        self._freqEMA = nn.ParameterList(nn.Parameter(torch.ones(m, ki) / ki, requires_grad=False) for ki in k)
        self._k = list(k)
        self._m = m
        self._ema = ema
        self._cdfs = None

    @property
    def NormalizedFreq(self) -> List[torch.Tensor]:
        return [(f / f.sum(-1, keepdim=True)).detach().clone() for f in self._freqEMA]

    @property
    def CDFs(self):
        """per level uint32 [m, k_l + 1], quantized to 16 bits (entropyCoder.py:50-63)."""
        def version(t):
            try:
                return t._version
            except RuntimeError:      # inference tensor: immutable
                return 0
        ver = tuple((version(f), f.data_ptr()) for f in self._freqEMA)
        if self._cdfs is None or self._cdfs[0] != ver:
            from .. import entropy
            import numpy as np
            tables = [np.stack([entropy.pmf_to_quantized_cdf(row.tolist()) for row in fr.cpu()]) for fr in self.NormalizedFreq]
            self._cdfs = (ver, tables)
        return self._cdfs[1]

    @torch.no_grad()
    def update(self, flat_hist: torch.Tensor):
        """EMA update from a flat int32 histogram [sum_l m*k_l] (already summed over ranks);
        same arithmetic as entropyCoder.py:38-43."""
        off = 0
        for lv, ki in enumerate(self._k):
            total = flat_hist[off:off + self._m * ki].reshape(self._m, ki).to(self._freqEMA[lv].dtype)
            off += self._m * ki
            normalized = total / total.sum(-1, keepdim=True)
            self._freqEMA[lv].copy_((1 - self._ema) * normalized + self._ema * self._freqEMA[lv])

    def _check(self, codes: List[torch.Tensor]):
        # same messages as entropyCoder.py:79-93
        info = "Please give codes with correct shape, for example, [[1, 2, 24, 24], [1, 2, 12, 12], ...], which is a `level` length list. each code has shape [n, m, h, w]. "
        if len(codes) < 1:
            raise RuntimeError("Length of codes is 0.")
        n, m = codes[0].shape[0], codes[0].shape[1]
        for code in codes:
            if n < 1:
                raise RuntimeError(info + "Now `n` = 0.")
            if code.shape[1] != m:
                raise RuntimeError(info + "Now `m` is inconsisitent.")
            if code.shape[0] != n:
                raise RuntimeError(info + "Now `n` is inconsisitent.")
        return n, m

    @torch.no_grad()
    def compress(self, codes: List[torch.Tensor]):
        """codes: L x int64 [n, m, h, w] -> (binaries: n lists of L byte strings, n CodeSize records)
        (entropyCoder.py:96-126; one batched, multi-threaded call per level instead of n Python-level calls)."""
        from .. import entropy
        n, m = self._check(codes)
        per_level = [entropy.encode_level(code, cdf) for code, cdf in zip(codes, self.CDFs)]
        binaries = [[per_level[lv][i] for lv in range(len(codes))] for i in range(n)]
        size = entropy.CodeSize([m] * len(codes), [c.shape[2] for c in codes], [c.shape[3] for c in codes], list(self._k))
        return binaries, [size for _ in range(n)]

    @torch.no_grad()
    def decompress(self, binaries, codeSizes) -> List[torch.Tensor]:
        """inverse of `compress` (entropyCoder.py:128-154): L x int64 [n, m, h, w] on this module's device."""
        from .. import entropy
        size = codeSizes[0]
        out = []
        for lv, cdf in enumerate(self.CDFs):
            streams = [b[lv] for b in binaries]
            out.append(entropy.decode_level(streams, size.m[lv], size.heights[lv], size.widths[lv], cdf)
                       .to(self._freqEMA[0].device))
        return out


class _multiCodebookQuantization(nn.Module):
    def __init__(self, codebook: nn.Parameter, freqEMA=None):
        super().__init__()
        self._m, self._k, self._d = codebook.shape
        self._codebook = codebook
        self._scale = math.sqrt(self._k)
        self._temperature = nn.Parameter(torch.ones((self._m, 1, 1, 1)))
        self._bound = LowerBound(_EPS)  # checkpoint key `_bound.bound` (quantizer.py:107)
        self._c2_cache = None

    def _tables(self):
        """Per codebook version: fp32 codebook, |c_k|^2 [m, k] (quantizer.py:165) and the split-fp16 packing the
        tensor-core kernel consumes."""
        try:
            cbv = self._codebook._version
        except RuntimeError:          # inference tensor (model built under torch.inference_mode()): immutable
            cbv = 0
        ver = (cbv, self._codebook.data_ptr())
        if self._c2_cache is None or self._c2_cache[0] != ver:
            from ..engine import pack_codebook
            with torch.no_grad():
                cb = self._codebook.detach().float().contiguous()
                self._c2_cache = (ver, cb, (cb ** 2).sum(-1).contiguous(), pack_codebook(cb))
        return self._c2_cache[1:]

    def _c2(self) -> torch.Tensor:
        return self._tables()[1]

    def _cb(self) -> torch.Tensor:
        return self._tables()[0]

    def encode_nhwc(self, x_f32: torch.Tensor, n: int, h: int, w: int, hist: Optional[torch.Tensor] = None,
                    engine: Optional[Engine] = None) -> torch.Tensor:
        cb, c2, packed = self._tables()
        return (engine or default_engine()).vq_assign(x_f32, cb, c2, n, h, w, hist=hist, packed=packed)

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """[n, c, h, w] fp32 -> int64 codes [n, m, h, w] (quantizer.py:144-150)."""
        n, c, h, w = x.shape
        if c != self._m * self._d:
            raise RuntimeError(f"expected {self._m * self._d} channels, got {c}")
        eng = default_engine()
        return self.encode_nhwc(eng.from_nchw(x, {"f32"}).f32, n, h, w, engine=eng)

    @torch.no_grad()
    def logits(self, x: torch.Tensor):
        """(code, logit [n,m,h,w,k]) -- the deterministic part of the soft path (quantizer.py:181-183,204):
        logit = -distance / sqrt(k) * max(temperature, Eps); code = argmin distance."""
        n, c, h, w = x.shape
        eng = default_engine()
        scale = self._bound(self._temperature.detach().float()).reshape(-1).contiguous()
        cb, c2, packed = self._tables()
        return eng.vq_assign(eng.from_nchw(x, {"f32"}).f32, cb, c2, n, h, w, logits=True, logit_scale=scale,
                             packed=packed)

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """(sample, code, oneHot, logit) as quantizer.py:232-239, forward values only (no autograd through the
        CUDA path).  The Gumbel noise is drawn by PyTorch (base.py:118-133); `_randomDrop` (quantizer.py:194-200)
        is a no-op here because UMGMQuantizer never gives this module a usable freqEMA (quantizer.py:399)."""
        _, logit = self.logits(x)
        eps = torch.finfo(logit.dtype).eps
        gumbels = -((-(torch.rand_like(logit).clamp_(eps, 1 - eps).log())).log())
        ySoft = (logit + gumbels).softmax(-1)
        sample = torch.zeros_like(logit).scatter_(-1, ySoft.max(-1, keepdim=True)[1], 1.0)
        code = logit.argmax(-1, keepdim=True)
        oneHot = torch.zeros_like(logit).scatter_(-1, code, 1)
        return sample, code[..., 0].contiguous(), oneHot, logit


class _multiCodebookDeQuantization(nn.Module):
    def __init__(self, codebook: nn.Parameter):
        super().__init__()
        self._m, self._k, self._d = codebook.shape
        self._codebook = codebook

    def _check(self, code: torch.Tensor):
        if code.dim() != 4 or code.shape[1] != self._m or code.dtype != torch.int64:
            raise RuntimeError(f"codes must be int64 [n, {self._m}, h, w], got {code.dtype} {tuple(code.shape)}")

    def decode_act(self, code: torch.Tensor, want, engine: Optional[Engine] = None,
                   status: Optional[torch.Tensor] = None) -> Act:
        self._check(code)
        return (engine or default_engine()).vq_dequant(code.contiguous(), self._codebook.detach().float().contiguous(),
                                                       want, status)

    @torch.no_grad()
    def decode(self, code: torch.Tensor) -> torch.Tensor:
        """int64 codes [n, m, h, w] -> [n, c, h, w] fp32 (quantizer.py:249-259)."""
        eng = default_engine()
        status = torch.zeros(1, dtype=torch.int32, device=code.device)
        out = eng.to_nchw(self.decode_act(code, {"f32"}, eng, status))
        if int(status.item()) != 0:
            raise RuntimeError(f"code index outside [0, {self._k})")
        return out

    @torch.no_grad()
    def forward(self, sample: torch.Tensor) -> torch.Tensor:
        """soft de-quantization of a one-hot / relaxed sample [n,m,h,w,k] (quantizer.py:262-274)."""
        n, m, h, w, k = sample.shape
        left = sample.reshape(n * m, h * w, k)
        right = self._codebook.detach().expand(n, m, k, self._d).reshape(n * m, k, self._d)
        return torch.bmm(left, right).reshape(n, m, h, w, self._d).permute(0, 1, 4, 2, 3).reshape(n, -1, h, w).contiguous()


class _quantizerEncoder(nn.Module):
    def __init__(self, quantizer, dequantizer, latentStageEncoder, quantizationHead, latentHead):
        super().__init__()
        self._quantizer = quantizer
        self._dequantizer = dequantizer
        self._latentStageEncoder = latentStageEncoder
        self._quantizationHead = quantizationHead
        self._latentHead = latentHead

    @property
    def Codebook(self):
        return self._quantizer._codebook

    def encode_act(self, eng: Engine, x: Act, next_needs, hist: Optional[torch.Tensor]):
        """quantizer.py:310-318 on engine activations: returns (residual Act or None, codes)."""
        z = eng.run_seq(self._latentStageEncoder, x, {"f32", "silu"})
        if self._latentHead is None:
            head = eng.run_seq(self._quantizationHead, z, {"f32"})
            return None, self._quantizer.encode_nhwc(head.f32, head.n, head.h, head.w, hist, eng)
        latent = list(self._latentHead)

        def quantize():
            head = eng.run_seq(self._quantizationHead, z, {"f32"})
            code = self._quantizer.encode_nhwc(head.f32, head.n, head.h, head.w, hist, eng)
            return code, self._dequantizer.decode_act(code, {"f32"}, eng)

        # all of latentHead but its last conv is independent of the code (quantizer.py:313-316)
        if eng.can_chain(z):
            # small maps: the layers of both heads are interleaved in one layer-chain launch, the VQ follows
            head, zl = eng.parallel(lambda: eng.run_seq(self._quantizationHead, z, {"f32"}),
                                    lambda: eng.run_seq(latent[:-1], z, eng.needs_of(latent[-1])), chain=True)
            code = self._quantizer.encode_nhwc(head.f32, head.n, head.h, head.w, hist, eng)
            deq = self._dequantizer.decode_act(code, {"f32"}, eng)
        else:
            (code, deq), zl = eng.parallel(quantize, lambda: eng.run_seq(latent[:-1], z, eng.needs_of(latent[-1])))
        return eng.run(latent[-1], zl, next_needs, tail=(deq.f32, -1.0)), code


class _quantizerDecoder(nn.Module):
    def __init__(self, dequantizer, dequantizationHead, sideHead, restoreHead):
        super().__init__()
        self._dequantizer = dequantizer
        self._dequantizationHead = dequantizationHead
        self._sideHead = sideHead
        self._restoreHead = restoreHead

    def decode_act(self, eng: Engine, code: torch.Tensor, former: Optional[Act], next_needs,
                   status: Optional[torch.Tensor]) -> Act:
        """quantizer.py:351-357 on engine activations."""
        q0 = self._dequantizer.decode_act(code, eng.needs_of(self._dequantizationHead[0]), eng, status)
        head_needs = eng.needs_of(self._restoreHead[0])
        if self._sideHead is not None:
            head = list(self._dequantizationHead)
            qh, side = eng.parallel(lambda: eng.run_seq(head[:-1], q0, eng.needs_of(head[-1])),
                                    lambda: eng.run_seq(self._sideHead, former, {"f32"}),
                                    chain=eng.can_chain(q0) and eng.can_chain(former))
            q = eng.run(head[-1], qh, head_needs, tail=(side.f32, 1.0))
        else:
            q = eng.run_seq(self._dequantizationHead, q0, head_needs)
        return eng.run_seq(self._restoreHead, q, next_needs)


class UMGMQuantizer(nn.Module):
    _components = ["latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead",
                   "restoreHead"]

    def __init__(self, channel: int, m: int, k: Union[int, List[int]], permutationRate: float,
                 components: Dict[str, Callable[[], nn.Module]]):
        super().__init__()
        if isinstance(k, int):
            k = [k]
        if channel % m != 0:
            raise ValueError(f"channel ({channel}) must be divisible by m ({m})")
        self._m, self._k = m, list(k)
        self._entropyCoder = CodeFrequency(m, self._k)
        fns = [components[key] for key in self._components]
        encoders, decoders = [], []
        for i, ki in enumerate(self._k):
            last = i == len(self._k) - 1
            latentStageEncoder, quantizationHead = fns[0](), fns[1]()
            latentHead = None if last else fns[2]()
            dequantizationHead = fns[3]()
            sideHead = None if last else fns[4]()
            restoreHead = fns[5]()
            # same init law as quantizer.py:398 ("SmallInit")
            codebook = nn.Parameter(nn.init.normal_(torch.empty(m, ki, channel // m),
                                                    std=math.sqrt(2 / (5 * channel / m))))
            quantizer = _multiCodebookQuantization(codebook)
            dequantizer = _multiCodebookDeQuantization(codebook)
            encoders.append(_quantizerEncoder(quantizer, dequantizer, latentStageEncoder, quantizationHead, latentHead))
            decoders.append(_quantizerDecoder(dequantizer, dequantizationHead, sideHead, restoreHead))
        self._encoders = nn.ModuleList(encoders)
        self._decoders = nn.ModuleList(decoders)

    @property
    def Codebooks(self):
        return [enc.Codebook for enc in self._encoders]

    @property
    def NormalizedFreq(self):
        return self._entropyCoder.NormalizedFreq

    @property
    def CDFs(self):
        return self._entropyCoder.CDFs

    def first_needs(self, eng: Engine):
        return eng.needs_of(self._encoders[0]._latentStageEncoder[0])

    def encode_act(self, eng: Engine, y: Act, hist: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """quantizer.py:411-420; `hist` (flat int32 [sum m*k_l]) is filled by the VQ kernel when given."""
        codes, off = [], 0
        x = y
        for lv, enc in enumerate(self._encoders):
            nxt = self.first_needs(eng) if lv + 1 < len(self._encoders) else None
            view = None if hist is None else hist[off:off + self._m * self._k[lv]]
            off += self._m * self._k[lv]
            x, code = enc.encode_act(eng, x, nxt, view)
            codes.append(code)
        return codes

    def decode_act(self, eng: Engine, codes: List[torch.Tensor], final_needs, status=None) -> Act:
        """quantizer.py:422-428."""
        if len(codes) != len(self._decoders):
            raise RuntimeError(f"expected {len(self._decoders)} code levels, got {len(codes)}")
        former = None
        for lv in reversed(range(len(self._decoders))):
            nxt = final_needs if lv == 0 else eng.needs_of(self._decoders[lv - 1]._sideHead[0])
            former = self._decoders[lv].decode_act(eng, codes[lv], former, nxt, status)
        return former

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> List[torch.Tensor]:
        eng = default_engine()
        codes = self.encode_act(eng, eng.from_nchw(x, self.first_needs(eng)))
        eng.flush()
        return codes

    @torch.no_grad()
    def decode(self, codes: List[torch.Tensor]) -> torch.Tensor:
        eng = default_engine()
        return eng.to_nchw(self.decode_act(eng, codes, {"f32"}))
