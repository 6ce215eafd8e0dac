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

    def __init__(self, m: Union[int, List[int]], k: List[int], ema: float = 0.9):
        """m: one codebook count for every level (EntropyCoder, entropyCoder.py:15-26, ema 0.9) or one per level
        (VariousMCoder, entropyCoder.py:293-303, ema 0.998)."""
        super().__init__()
        self._k = list(k)
        self._m = m
        self._freqEMA = nn.ParameterList(nn.Parameter(torch.ones(mi, ki) / ki, requires_grad=False)
                                         for mi, ki in zip(self.level_m(), self._k))
        self._ema = ema
        self._cdfs = None

    def level_m(self) -> List[int]:
        return list(self._m) if isinstance(self._m, (list, tuple)) else [self._m] * len(self._k)

    def hist_size(self) -> int:
        """length of the flat int32 code histogram: sum over levels of m_l * k_l (level-major, the order of _freqEMA)"""
        return sum(mi * ki for mi, ki in zip(self.level_m(), self._k))

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
        for lv, (mi, ki) in enumerate(zip(self.level_m(), self._k)):
            total = flat_hist[off:off + mi * ki].reshape(mi, ki).to(self._freqEMA[lv].dtype)
            off += mi * ki
            normalized = total / total.sum(-1, keepdim=True)
            self._freqEMA[lv].copy_((1 - self._ema) * normalized + self._ema * self._freqEMA[lv])

    @torch.no_grad()
    def forward(self, oneHotCodes: List[torch.Tensor]):
        """entropyCoder.py:28-44 / :306-322: per level, count = one-hot.sum((0, 2, 3)) -> all-reduce over the ranks (when a
        process group is up) -> normalise -> EMA.  (On the inference path the same counts come out of the VQ launch as
        an int32 histogram, see `update`.)"""
        import torch.distributed as dist
        for lv, code in enumerate(oneHotCodes):
            total = code.sum((0, 2, 3))
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(total)
            normalized = total / total.sum(-1, keepdim=True)
            self._freqEMA[lv].copy_((1 - self._ema) * normalized + self._ema * self._freqEMA[lv])
        self._cdfs = None

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
        if len(binaries) < 1 or len(binaries) != len(codeSizes):
            raise RuntimeError("decompress: one CodeSize record per image is required")
        size = codeSizes[0]
        levels = len(self._k)
        # the header comes from a file: everything in it is checked against the model before the coder sees it
        for cs in codeSizes:
            if (list(cs.m), list(cs.heights), list(cs.widths), list(cs.k)) != \
                    (list(size.m), list(size.heights), list(size.widths), list(size.k)):
                raise RuntimeError("decompress: all images of a batch must share one CodeSize")
        if not (len(size.m) == len(size.heights) == len(size.widths) == len(size.k) == levels):
            raise RuntimeError(f"decompress: header describes {len(size.k)} code levels, the model has {levels}")
        if list(size.k) != list(self._k) or list(size.m) != self.level_m():
            raise RuntimeError(f"decompress: header (m = {list(size.m)}, k = {list(size.k)}) does not match the model "
                               f"(m = {self.level_m()}, k = {list(self._k)})")
        for b in binaries:
            if len(b) != levels:
                raise RuntimeError(f"decompress: expected {levels} streams per image, got {len(b)}")
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
        # quantizer.py:109: whatever the owner hands over is stored; only a Parameter (ResidualBackwardQuantizer passes
        # the entropy coder's, :618) shows up in the state_dict -- UMGMQuantizer passes a float (:399)
        if isinstance(freqEMA, nn.Parameter):
            self._freqEMA = freqEMA
        self._c2_cache = None

    @torch.no_grad()
    def reAssignCodebook(self, freq: torch.Tensor) -> torch.Tensor:
        """quantizer.py:111-136 (codebook maintenance, SURVEY.md 8f NEXT-4): per codebook, codewords whose normalised
        frequency is below Eps are overwritten by the most frequently used ones; when more than half were never
        used, a random half of the never-used ones (torch.randperm, like upstream) is kept untouched.  Returns the
        flat [m*k] bool mask of codewords that moved by more than 1e-4 (squared L2).  Control plane: plain torch ops
        on the codebook's device."""
        new = self._codebook.detach().clone()
        freq = freq.to(self._codebook.device).detach().clone()
        half = self._k // 2
        for j in range(self._m):
            row = freq[j]
            never = row < _EPS
            count = int(never.sum())
            if count > half:
                marks = torch.zeros((count,), device=self._codebook.device)
                marks[torch.randperm(count)[half:]] = -1.0     # these never-used slots are left alone this time
                row[never] = marks
                never = (row < _EPS) * (row > -_EPS)
                count = int(never.sum())
            order = torch.argsort(row, descending=True)
            new[j, never] = self._codebook.detach()[j][order][:count]
        moved = ((new - self._codebook.detach()) ** 2).sum(-1) > 1e-4
        self._codebook.copy_(new)        # in place under no_grad: bumps the version counter the weight caches key on
        return moved.flatten()

    @torch.no_grad()
    def syncCodebook(self):
        """quantizer.py:138-142: rank 0's codebook wins (broadcast of [m, k, d] over NCCL / gloo)."""
        import torch.distributed as dist
        codebook = self._codebook.detach().clone()
        dist.broadcast(codebook, 0)
        self._codebook.copy_(codebook)

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

    def _randomDrop(self, logit: torch.Tensor) -> torch.Tensor:
        """quantizer.py:194-200: frequently used codewords are masked out at random (logit -= 1e9) so that rare ones get
        picked; the exponent falls from `bits` (no code used) to 1 (all codes used)."""
        bits = math.log2(self._k)
        freq = self._freqEMA.detach().to(logit.device)
        usage = (freq > _EPS).float().mean().clamp(0.0, 1.0)
        mask = (torch.rand_like(logit) ** (-(bits - 1) * (usage ** 2) + bits)) < freq[:, None, None, ...]
        # upstream: `logit[randomMask] += -1e9` (quantizer.py:199).  Same values, but as a select instead of a boolean-mask
        # scatter: the latter needs the number of set bits on the host, i.e. a device synchronisation per level, which also
        # keeps a training step from being captured in a CUDA graph
        return torch.where(mask, logit + (-1e9), logit)

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """(sample, code, oneHot, logit) as quantizer.py:202-239 -- FORWARD VALUES ONLY (no autograd through the CUDA
        path; training is SURVEY 8f NEXT-3).  distance / logits: the VQ launch; `_randomDrop` (only when the owner
        registered a frequency Parameter, i.e. ResidualBackwardQuantizer -- UMGMQuantizer hands over a float upstream,
        quantizer.py:399, and its `_randomDrop` would raise) and the hard Gumbel sample `y_hard - y_soft + y_soft`
        (base.py:118-133) are the reference's own torch expressions, with PyTorch's RNG, on the logits' device."""
        _, logit = self.logits(x)
        if isinstance(getattr(self, "_freqEMA", None), torch.Tensor):
            logit = self._randomDrop(logit)
        eps = torch.finfo(logit.dtype).eps
        gumbels = -((-(torch.rand_like(logit).clamp_(eps, 1 - eps).log())).log())
        ySoft = (logit + gumbels).softmax(-1)
        yHard = torch.zeros_like(logit).scatter_(-1, ySoft.max(-1, keepdim=True)[1], 1.0)
        sample = yHard - ySoft + ySoft
        code = logit.argmax(-1, keepdim=True)
        oneHot = torch.zeros_like(logit).scatter_(-1, code, 1).contiguous()
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

    def reAssignCodebook(self, freq: torch.Tensor) -> torch.Tensor:     # quantizer.py:291-292
        return self._quantizer.reAssignCodebook(freq)

    def syncCodebook(self):                                             # quantizer.py:294-295
        return self._quantizer.syncCodebook()

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """quantizer.py:320-328, forward values: (q, residual or None, code, oneHot, logit) on NCHW tensors"""
        eng = default_engine()
        z = eng.run_seq_nchw(self._latentStageEncoder, x)
        q, code, oneHot, logit = self._quantizer(eng.run_seq_nchw(self._quantizationHead, z))
        if self._latentHead is None:
            return q, None, code, oneHot, logit
        return q, eng.run_seq_nchw(self._latentHead, z) - self._dequantizer(q), code, oneHot, logit

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

    @torch.no_grad()
    def forward(self, q: torch.Tensor, formerLevel: Optional[torch.Tensor]):
        """quantizer.py:359-365, forward values: q is the (relaxed) one-hot sample [n, m, h, w, k]"""
        eng = default_engine()
        xHat = eng.run_seq_nchw(self._dequantizationHead, self._dequantizer(q))
        if self._sideHead is not None:
            xHat = xHat + eng.run_seq_nchw(self._sideHead, formerLevel)
        return eng.run_seq_nchw(self._restoreHead, xHat)

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

    def hist_size(self) -> int:
        return self._entropyCoder.hist_size()

    def reAssignCodebook(self) -> torch.Tensor:
        """quantizer.py:430-436: fraction of codewords re-assigned over all levels"""
        moved = [enc.reAssignCodebook(freq) for enc, freq in zip(self._encoders, self.NormalizedFreq)]
        return torch.cat(moved).float().mean()

    def syncCodebook(self):
        """quantizer.py:438-441"""
        import torch.distributed as dist
        dist.barrier()
        for enc in self._encoders:
            enc.syncCodebook()

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
    def forward(self, x: torch.Tensor):
        """quantizer.py:443-467, FORWARD VALUES ONLY: (yHat, codes, logits); updates the frequency EMA from the one-hot
        codes like upstream (:464).  (Upstream this method cannot run inside `Compressor`: the quantization modules get
        a float where `_randomDrop` expects the frequency tensor, quantizer.py:399 / SURVEY 8a row a14; here the drop is
        skipped in that case.)"""
        quantizeds, codes, oneHots, logits = [], [], [], []
        for enc in self._encoders:
            quantized, x, code, oneHot, logit = enc(x)
            quantizeds.append(quantized)
            codes.append(code)
            oneHots.append(oneHot)
            logits.append(logit)
        former = None
        for dec, quantized in zip(self._decoders[::-1], quantizeds[::-1]):
            former = dec(quantized, former)
        self._entropyCoder(oneHots)
        return former, codes, logits

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


class ResidualBackwardQuantizer(nn.Module):
    """The `Neon` tokenizer's quantizer (mcquic/modules/quantizer.py:577-765, on VariousMQuantizer :88-93): every level
    has m = 1 and shares ONE codebook [1, k, 8]; `size` lists the latent grid per level and must halve or stay equal
    from left to right (:600-657).  Encoding first produces all latents (large to small), then codes the residual of
    each latent against the up-projection (`_backwards`) of what the smaller levels already explain, from the smallest
    level back to the largest -- codes are therefore returned SMALLEST level first (:675-693); decoding sums the
    de-quantised codes into the running reconstruction level by level (:695-703).

    State-dict keys as upstream: `_entropyCoder._freqEMA.{j}`, `_encoders.{l}.{0..3}`, `_backwards.{l}.{0..3}`,
    `_decoders.{l}.{0..3}`, `_quantizers.{l}.{_codebook,_temperature,_freqEMA,_bound.bound}`,
    `_dequantizers.{l}._codebook` (all codebook keys alias one Parameter; `_quantizers.{l}._freqEMA` aliases
    `_entropyCoder._freqEMA.{L-1-l}`)."""

    channel = 8

    def __init__(self, k: int, size: List[int], denseNorm: bool):
        super().__init__()
        from ..nn import AttentionBlock, ResidualBlock, ResidualBlockShuffle, ResidualBlockWithStride, conv1x1
        c = self.channel
        self._m, self._k = [1] * len(size), [k] * len(size)
        self._entropyCoder = CodeFrequency(self._m, self._k, ema=0.998)
        codebook = nn.Parameter(nn.init.trunc_normal_(torch.empty(1, k, c), std=math.sqrt(2 / (5 * c))))
        encoders, backwards, decoders, quantizers, dequantizers = [], [], [], [], []
        self._strided: List[bool] = []
        last = size[0] * 2
        for i, this in enumerate(size):
            if this == last // 2:
                strided = True
            elif this == last:
                strided = False
            else:
                raise ValueError("The given size sequence does not half or equal to from left to right.")
            self._strided.append(strided)

            def up():
                return nn.Sequential(conv1x1(c, c * 4, bias=False),
                                     ResidualBlockShuffle(c * 4, c * 4, 2, 1, denseNorm) if strided
                                     else ResidualBlock(c * 4, c * 4, 1, denseNorm),
                                     AttentionBlock(c * 4, 1, denseNorm), ResidualBlock(c * 4, c, 1, denseNorm))

            encoders.append(nn.Sequential(ResidualBlock(c, c * 4, 1, denseNorm), AttentionBlock(c * 4, 1, denseNorm),
                                          ResidualBlockWithStride(c * 4, c * 4, 2, 1, denseNorm) if strided
                                          else ResidualBlock(c * 4, c * 4, 1, denseNorm),
                                          conv1x1(c * 4, c, bias=False)))
            quantizers.append(_multiCodebookQuantization(codebook, self._entropyCoder._freqEMA[-(i + 1)]))
            dequantizers.append(_multiCodebookDeQuantization(codebook))
            backwards.append(up() if i < len(size) - 1 else nn.Identity())
            decoders.append(up())
            last = this
        self._encoders = nn.ModuleList(encoders)
        self._decoders = nn.ModuleList(decoders)
        self._backwards = nn.ModuleList(backwards)
        self._quantizers = nn.ModuleList(quantizers)
        self._dequantizers = nn.ModuleList(dequantizers)

    @property
    def Codebooks(self):
        return [q._codebook for q in self._quantizers]

    @property
    def NormalizedFreq(self):
        return self._entropyCoder.NormalizedFreq

    @property
    def CDFs(self):
        return self._entropyCoder.CDFs

    def hist_size(self) -> int:
        return self._entropyCoder.hist_size()

    def reAssignCodebook(self) -> torch.Tensor:                        # quantizer.py:714-720
        moved = [q.reAssignCodebook(freq) for q, freq in zip(self._quantizers, self.NormalizedFreq)]
        return torch.cat(moved).float().mean()

    def syncCodebook(self):                                            # quantizer.py:722-725
        import torch.distributed as dist
        dist.barrier()
        for q in self._quantizers:
            q.syncCodebook()

    def first_needs(self, eng: Engine):
        return eng.needs_of(self._encoders[0][0])

    def _up_act(self, eng: Engine, net: nn.Sequential, q: Act, want) -> Act:
        return eng.run_seq(net, q, want)

    def encode_act(self, eng: Engine, y: Act, hist: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """quantizer.py:675-693.  hist: flat int32 [L*k]; segment j counts codes[j] (the order of `_freqEMA`, :616)."""
        levels = len(self._encoders)
        latents, x = [], y
        for lv, enc in enumerate(self._encoders):
            want = {"f32"} | (eng.needs_of(self._encoders[lv + 1][0]) if lv + 1 < levels else set())
            x = eng.run_seq(enc, x, want)
            latents.append(x)
        codes: List[torch.Tensor] = []
        current: Optional[Act] = None
        k = self._k[0]
        for j, lv in enumerate(reversed(range(levels))):
            lat = latents[lv]
            residual = lat.f32 if current is None else eng.add_scaled(lat.f32, current.f32, -1.0,
                                                                      (lat.n, lat.h, lat.w, lat.c), {"f32"}).f32
            view = None if hist is None else hist[j * k:(j + 1) * k]
            code = self._quantizers[lv].encode_nhwc(residual, lat.n, lat.h, lat.w, view, eng)
            codes.append(code)
            if lv == 0:
                break                                    # upstream still up-projects the last level; nothing reads it
            back = self._backwards[lv]
            if isinstance(back, nn.Identity):
                current = self._dequantizers[lv].decode_act(code, {"f32"}, eng)
            else:
                q = self._dequantizers[lv].decode_act(code, eng.needs_of(back[0]), eng)
                current = eng.run_seq(back, q, {"f32"})
        return codes

    def decode_act(self, eng: Engine, codes: List[torch.Tensor], final_needs, status=None) -> Act:
        """quantizer.py:695-703 (codes smallest level first)."""
        levels = len(self._decoders)
        if len(codes) != levels:
            raise RuntimeError(f"expected {levels} code levels, got {len(codes)}")
        former: Optional[Act] = None
        for lv, code in zip(reversed(range(levels)), codes):
            dec = self._decoders[lv]
            needs = eng.needs_of(dec[0])
            if former is None:
                q = self._dequantizers[lv].decode_act(code, needs, eng, status)
            else:
                q0 = self._dequantizers[lv].decode_act(code, {"f32"}, eng, status)
                if (q0.n, q0.h, q0.w, q0.c) != (former.n, former.h, former.w, former.c):
                    raise RuntimeError(f"code level of grid {q0.h}x{q0.w} does not continue a {former.h}x{former.w} "
                                       "reconstruction")
                q = eng.add_scaled(q0.f32, former.f32, 1.0, (q0.n, q0.h, q0.w, q0.c), needs)
            former = eng.run_seq(dec, q, final_needs if lv == 0 else {"f32"})
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

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """quantizer.py:727-765, FORWARD VALUES ONLY: (restored latent, codes, logits), codes / logits smallest level first;
        the entropy coder's frequency EMA is updated from the one-hot codes like upstream (:763)."""
        eng = default_engine()
        latents = []
        for enc in self._encoders:
            x = eng.run_seq_nchw(enc, x)
            latents.append(x)
        quantizeds, codes, oneHots, logits = [], [], [], []
        current = torch.zeros_like(latents[-1])
        for lv in reversed(range(len(latents))):
            sample, code, oneHot, logit = self._quantizers[lv](latents[lv] - current)
            quantized = self._dequantizers[lv](sample)
            quantizeds.append(quantized)
            codes.append(code)
            oneHots.append(oneHot)
            logits.append(logit)
            back = self._backwards[lv]
            current = quantized if isinstance(back, nn.Identity) else eng.run_seq_nchw(back, quantized)
        former = torch.zeros_like(quantizeds[0])
        for lv, quantized in zip(reversed(range(len(latents))), quantizeds):
            former = eng.run_seq_nchw(self._decoders[lv], former + quantized)
        self._entropyCoder(oneHots)
        return former, codes, logits

    @torch.no_grad()
    def residual_backward(self, code: torch.Tensor, level: int) -> torch.Tensor:
        """quantizer.py:670-673: up-projection of one level's codes, indexed from the END like upstream
        (`self._dequantizers[-level]`)."""
        eng = default_engine()
        dq, back = self._dequantizers[-level], self._backwards[-level]
        if isinstance(back, nn.Identity):
            return dq.decode(code)
        return eng.to_nchw(eng.run_seq(back, dq.decode_act(code, eng.needs_of(back[0]), eng), {"f32"}))

    @torch.no_grad()
    def residual_forward(self, code: torch.Tensor, formerLevel: Optional[torch.Tensor], level: int) -> torch.Tensor:
        """quantizer.py:705-712"""
        if formerLevel is None and level > 0:
            raise RuntimeError("For reconstruction after level-0, you should provide not None formerLevel as input.")
        if formerLevel is not None and level == 0:
            raise RuntimeError("For reconstruction at level-0, you should provide None formerLevel as input.")
        eng = default_engine()
        dec, dq = self._decoders[-(level + 1)], self._dequantizers[-(level + 1)]
        needs = eng.needs_of(dec[0])
        if formerLevel is None:
            q = dq.decode_act(code, needs, eng)
        else:
            q0 = dq.decode_act(code, {"f32"}, eng)
            former = eng.from_nchw(formerLevel, {"f32"})
            q = eng.add_scaled(q0.f32, former.f32, 1.0, (q0.n, q0.h, q0.w, q0.c), needs)
        return eng.to_nchw(eng.run_seq(dec, q, {"f32"}))
