"""`Compressor` with the reference's constructor, state_dict layout and encode/decode contract
(mcquic/modules/compressor.py:18-117 BaseCompressor, :120-177 Compressor), executed by the CUDA engine.

    Compressor(channel, m, k).to("cuda").eval().load_state_dict(reference_checkpoint["model"])
    codes = model.encode(x)        # x fp32 [n,3,H,W] in [-1,1]  ->  L x int64 [n, m, h_l, w_l]
    xHat  = model.decode(codes)    # fp32 [n, 3, H_pad, W_pad]
"""
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch
from torch import nn

from ..engine import Act, Engine
from ..nn import AttentionBlock, ResidualBlock, ResidualBlockShuffle, ResidualBlockWithStride, conv3x3, pixelShuffle3x3
from .quantizer import ResidualBackwardQuantizer, UMGMQuantizer

ALIGN_BASE = 128  # mcquic/data/transforms.py:82


def aligned_pad_amounts(h: int, w: int, base: int = ALIGN_BASE) -> Tuple[int, int, int, int]:
    """(top, left, padded_h, padded_w) exactly as AlignedPadding.forward splits them (transforms.py:86-99)."""
    wPadding = ((w // base + 1) * base - w) % base
    hPadding = ((h // base + 1) * base - h) % base
    return hPadding // 2, wPadding // 2, h + hPadding, w + wPadding


class _ShapeCache:
    """LRU of per-input-shape CUDA graphs / host-I/O pipelines (static buffers + captured activations: hundreds of MB per
    shape at batch 64).  Bounded by entry count and by the device bytes the entries hold, so that a service fed
    arbitrary image sizes does not grow without limit; evicted graphs are re-captured on their next use."""

    def __init__(self, max_entries: int = 12, max_bytes: int = 24 << 30):
        self.max_entries, self.max_bytes = max_entries, max_bytes
        self._d: "OrderedDict[object, Tuple[object, int]]" = OrderedDict()

    def get(self, key):
        hit = self._d.get(key)
        if hit is None:
            return None
        self._d.move_to_end(key)
        return hit[0]

    def put(self, key, value, nbytes: int):
        self._d[key] = (value, max(0, int(nbytes)))
        self._d.move_to_end(key)
        while len(self._d) > 1 and (len(self._d) > self.max_entries or self.bytes() > self.max_bytes):
            self._d.popitem(last=False)

    def bytes(self) -> int:
        return sum(b for _, b in self._d.values())

    def clear(self):
        self._d.clear()

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        return key in self._d


class BaseCompressor(nn.Module):
    def __init__(self, encoder: nn.Module, quantizer: UMGMQuantizer, decoder: nn.Module):
        super().__init__()
        self._encoder = encoder
        self._decoder = decoder
        self._quantizer = quantizer
        self._qp = "-1"
        self._engine: Optional[Engine] = None
        self.encode_passes = 3  # split-fp16 x3: fp32-grade, code indices match the fp32 reference
        self.decode_passes = 1  # single fp16 pass: TF32-grade, pixels within 1e-3
        # encode/decode are ~170 dependent launches each: replay them as one CUDA graph per input shape
        self.use_graphs = True
        self._graphs = _ShapeCache()
        self._weights_seen = None  # fingerprint of the parameters the cached graphs / packed weights were built from
        self.graph_launches = 0  # kernels launched through graph replays (the library counts eager launches)
        # host I/O pipeline: a pinned host batch is processed in up to 4 slices through the first (encode) / last
        # (decode) full-resolution layers so that the PCIe copies overlap the convolutions
        self._pipes = _ShapeCache()
        self._copy_stream = None
        # L2-resident sub-batches: the full-resolution part of the analysis / synthesis transform is run per slice of this
        # many images (0 = whole batch at once), so that a layer's output is still in the 126 MB L2 when the next layer of
        # the same slice reads it (at batch 64 one 64x64x128 activation is 134 MB fp32 + 67 MB planes: every layer streams
        # through HBM).  Results are bit-identical (images are independent).
        # Measured at batch 64 (tools/exp_slices.py, profiles/r2_decode_slices_experiment.txt): with round 1's drain 32-image
        # decode slices won (5.51 -> 5.37 ms); with the bulk-store drain the whole batch at once does (4.58 vs 4.81 ms) -- the
        # layers no longer wait on their own global stores, and half-size launches have twice the kernel boundaries.
        # Encode always preferred the whole batch (7.37 vs 7.62 ms at 32).
        self.encode_slice = 0
        self.decode_slice = 0

    @property
    def QuantizationParameter(self) -> str:
        return self._qp

    @QuantizationParameter.setter
    def QuantizationParameter(self, qp: str):
        self._qp = qp

    @property
    def Codebooks(self):
        return self._quantizer.Codebooks

    @property
    def NormalizedFreq(self):
        return self._quantizer.NormalizedFreq

    @property
    def CodeUsage(self):
        """compressor.py:62-64: fraction of codewords with a non-negligible frequency"""
        return torch.cat([(freq > 1e-6).flatten() for freq in self._quantizer.NormalizedFreq]).float().mean()

    def reAssignCodebook(self) -> torch.Tensor:
        """compressor.py:45-46 (codebook maintenance between epochs); drops the captured graphs' packed codebooks"""
        out = self._quantizer.reAssignCodebook()
        self.invalidate()
        return out

    def syncCodebook(self):
        """compressor.py:48-49"""
        out = self._quantizer.syncCodebook()
        self.invalidate()
        return out

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine()
        return self._engine

    def set_impl(self, impl: str):
        """'tcgen05' (default) or 'simt' (fp32 CUDA-core cross-check kernels)."""
        self._engine = Engine(impl)

    def _check_image(self, x: torch.Tensor):
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an image batch [n, 3, h, w], got {tuple(x.shape)}")
        if x.shape[0] == 0:      # upstream raises too (quantizer.py:158: reshape of 0 elements with -1 is ambiguous)
            raise RuntimeError("cannot encode an empty batch")
        if not x.is_cuda and not self.engine.emulated and not self._host_batch_ok(x):
            raise RuntimeError("mcquic_b200 runs on CUDA tensors (or pinned fp32 host batches a CUDA-resident model "
                               "streams in); there is no CPU fallback")

    @staticmethod
    def host_slices(n: int, small_first: bool, itemsize: int = 4) -> List[Tuple[int, int]]:
        """Batch slices [(n0, n1), ...] a host batch of n images is streamed in.  Up to four slices of >= 8 images keep
        the per-slice layers efficient (measured: eight equal slices cost in extra dependent launches what the shorter
        exposed copy saves); the slice whose PCIe copy cannot overlap anything -- the first one going in, the last one
        coming out -- is split once more into a quarter and the rest, so only ~n/16 images' worth of copy stays exposed."""
        if itemsize == 1:
            # uint8 images: a quarter of the bytes -- the copy of half a batch (6 MB at 64 x 3 x 256 x 256, ~0.12 ms) is all
            # that stays exposed with two slices, and every further slice costs more in launches / tile quantisation of the
            # per-slice layers than it hides (measured: the 5-slice schedule left e2e 1.0 ms above the device-resident step
            # even with the bytes cut by 4)
            if n % 2 == 0 and n // 2 >= 8:
                return [(0, n // 2), (n // 2, n)]
            return [(0, n)]
        ch = next((c for c in (4, 2) if n % c == 0 and n // c >= 8), 1)
        nc = n // ch
        bounds = [(c * nc, (c + 1) * nc) for c in range(ch)]
        if ch > 1 and nc % 4 == 0:
            if small_first:
                bounds = [(0, nc // 4), (nc // 4, nc)] + bounds[1:]
            else:
                bounds = bounds[:-1] + [(n - nc, n - nc // 4), (n - nc // 4, n)]
        return bounds

    def _device(self) -> torch.device:
        return self._encoder[0].weight.device

    def _host_batch_ok(self, t: torch.Tensor) -> bool:
        """A pinned, contiguous host batch (fp32 in [-1, 1], or uint8 images as demo.compressImage / decompressImage
        take and return them) that the chunked copy/compute pipeline can stream."""
        return (not t.is_cuda and t.is_pinned() and t.dtype in (torch.float32, torch.uint8) and t.is_contiguous()
                and t.dim() == 4
                and self.use_graphs and not self.engine.emulated and self._device().type == "cuda"
                and isinstance(self._encoder[1], ResidualBlock) and isinstance(self._decoder[5], ResidualBlock))

    def invalidate(self):
        """Drop captured graphs and repacked weights.  Called automatically when the parameters change (see
        `_check_weights`); only writes that bypass autograd's version counters (`p.data.copy_()`) need an explicit call."""
        self._graphs.clear()
        self._pipes.clear()
        self._weights_seen = None
        if self._engine is not None:
            self._engine._packed.clear()

    def _weights_fingerprint(self) -> int:
        """cheap identity of the current parameter values: the sum of the in-place version counters of every parameter
        and buffer -- an optimizer step, `load_state_dict`, `p.copy_()` all bump one (`.to()` / `load_state_dict` also go
        through `invalidate`)"""
        ts = self.__dict__.get("_fp_tensors")
        if ts is None:
            ts = []
            for t in list(self.parameters()) + list(self.buffers()):
                try:
                    t._version
                    ts.append(t)
                except RuntimeError:      # inference tensor: immutable, nothing to track
                    pass
            self.__dict__["_fp_tensors"] = ts
        return sum(t._version for t in ts) + (len(ts) << 40)

    def _check_weights(self):
        """graphs replay launches whose operands (packed weights, packed codebooks) were baked in at capture time: before
        any replay make sure the parameters are still the ones they were captured from"""
        fp = self._weights_fingerprint()
        if fp != self._weights_seen:
            if self._weights_seen is not None:
                self._graphs.clear()
                self._pipes.clear()
            self._weights_seen = fp
            if not self.engine.emulated and self._device().type == "cuda":
                self.engine.prepare(self)       # all layers repacked with one host sync instead of one per layer

    def load_state_dict(self, *args, **kwargs):
        self.invalidate()
        self.__dict__.pop("_fp_tensors", None)
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        self.__dict__.pop("_fp_tensors", None)
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ eager bodies
    def _encode_eager(self, x: torch.Tensor, hist: Optional[torch.Tensor]) -> List[torch.Tensor]:
        eng = self.engine
        eng.passes = self.encode_passes
        n, _, h, w = x.shape
        sl = self.encode_slice
        if sl and n > sl and isinstance(self._encoder[-1], ResidualBlock):
            _, _, hp, wp = aligned_pad_amounts(h, w)
            y = eng.alloc_act(n, hp // 8, wp // 8, self._encoder[0].out_channels, self._quantizer.first_needs(eng), x.device)
            mods = list(self._encoder)
            for n0 in range(0, n, sl):
                n1 = min(n, n0 + sl)
                y0 = eng.stem(mods[0], x[n0:n1], aligned_pad_amounts(h, w), eng.needs_of(mods[1]))
                t = eng.run_seq(mods[1:-1], y0, eng.needs_of(mods[-1]))
                eng.run(mods[-1], t, self._quantizer.first_needs(eng), into=y.batch_slice(n0, n1))
        else:
            y0 = eng.stem(self._encoder[0], x, aligned_pad_amounts(h, w), eng.needs_of(self._encoder[1]))
            y = eng.run_seq(list(self._encoder)[1:], y0, self._quantizer.first_needs(eng))
        codes = self._quantizer.encode_act(eng, y, hist)
        eng.flush()
        return codes

    def _decode_eager(self, codes: List[torch.Tensor], status: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        eng.passes = self.decode_passes
        yHat = self._quantizer.decode_act(eng, codes, eng.needs_of(self._decoder[0]), status)
        sl = self.decode_slice
        n = yHat.n
        if sl and n > sl:
            mods = list(self._decoder)
            y0 = eng.run(mods[0], yHat, eng.needs_of(mods[1]))
            out = torch.empty((n, 3, yHat.h * 8, yHat.w * 8), dtype=torch.float32, device=yHat.f32.device
                              if yHat.f32 is not None else codes[0].device)
            for n0 in range(0, n, sl):
                n1 = min(n, n0 + sl)
                t = eng.run_seq(mods[1:-1], y0.batch_slice(n0, n1), eng.needs_of(mods[-1]))
                eng.run(mods[-1], t, set(), into=Act(n1 - n0, out.shape[2], out.shape[3], 3, f32=out[n0:n1]))
        else:
            out = eng.run_seq(list(self._decoder), yHat, set()).f32
        eng.flush()
        return out

    def _graph(self, key, make_static, body, cache: bool = True):
        """Capture `body(*static)` once per key; returns (graph, static inputs, static outputs, #launches).
        cache=False: the caller (a host-I/O pipeline, itself an LRU entry) owns the graph."""
        entry = self._graphs.get(key) if cache else None
        if entry is None:
            from .. import _lib
            mem0 = torch.cuda.memory_allocated()
            static = make_static()
            cur = torch.cuda.current_stream()
            warm = torch.cuda.Stream()
            warm.wait_stream(cur)
            with torch.cuda.stream(warm):       # eager warm-up: repacks weights, sets kernel attributes
                body(*static)
            cur.wait_stream(warm)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph):
                out = body(*static)
            entry = (graph, static, out, _lib.launch_count() - before)
            if cache:
                self._graphs.put(key, entry, torch.cuda.memory_allocated() - mem0)
        return entry

    # ------------------------------------------------------------------ host I/O pipeline
    def _encode_head(self, x_chunk: torch.Tensor, into: Act):
        """stem + first ResidualBlock (full resolution) on a batch slice, written into the full-batch activation."""
        eng = self.engine
        eng.passes = self.encode_passes
        _, _, h, w = x_chunk.shape
        y0 = eng.stem(self._encoder[0], x_chunk, aligned_pad_amounts(h, w), eng.needs_of(self._encoder[1]))
        eng.run(self._encoder[1], y0, eng.needs_of(self._encoder[2]), into=into)
        eng.flush()

    def _encode_tail(self, y1: Act, hist: Optional[torch.Tensor]) -> List[torch.Tensor]:
        eng = self.engine
        eng.passes = self.encode_passes
        y = eng.run_seq(list(self._encoder)[2:], y1, self._quantizer.first_needs(eng))
        codes = self._quantizer.encode_act(eng, y, hist)
        eng.flush()
        return codes

    def _decode_main(self, codes: List[torch.Tensor], status: torch.Tensor) -> Act:
        eng = self.engine
        eng.passes = self.decode_passes
        yHat = self._quantizer.decode_act(eng, codes, eng.needs_of(self._decoder[0]), status)
        y = eng.run_seq(list(self._decoder)[:5], yHat, eng.needs_of(self._decoder[5]))
        eng.flush()
        return y

    def _decode_tail(self, y4_chunk: Act, out_chunk: torch.Tensor):
        """last ResidualBlock + final pixel-shuffle conv on a batch slice -> NCHW pixels of that slice."""
        eng = self.engine
        eng.passes = self.decode_passes
        y5 = eng.run(self._decoder[5], y4_chunk, eng.needs_of(self._decoder[6]))
        n, c, h, w = out_chunk.shape
        eng.run(self._decoder[6], y5, set(), into=Act(n, h, w, c, f32=out_chunk))
        eng.flush()

    def _streams(self, dev):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        return torch.cuda.current_stream(dev), self._copy_stream

    def _encode_pipelined(self, x: torch.Tensor, hist: Optional[torch.Tensor]) -> List[torch.Tensor]:
        """x: pinned host batch.  Chunk c is copied on the copy stream while chunk c-1 runs stem + first block."""
        dev = self._device()
        n, _, h, w = x.shape
        bounds = self.host_slices(n, small_first=True, itemsize=x.element_size())
        key = ("enc", tuple(x.shape), x.dtype, self.encode_passes, dev)
        pipe = self._pipes.get(key)
        with torch.cuda.device(dev):
            if pipe is None:
                mem0 = torch.cuda.memory_allocated()
                eng = self.engine
                eng.passes = self.encode_passes
                total = self._quantizer.hist_size()
                sx = torch.empty(tuple(x.shape), dtype=x.dtype, device=dev)
                sh = torch.zeros(total, dtype=torch.int32, device=dev)
                _, _, hp, wp = aligned_pad_amounts(h, w)
                y1 = eng.alloc_act(n, hp // 2, wp // 2, self._encoder[0].out_channels, eng.needs_of(self._encoder[2]), dev)
                sx.zero_()
                heads, launches = [], 0
                for c, (n0, n1) in enumerate(bounds):
                    g, _, _, l = self._graph(key + ("head", c), list,
                                             lambda n0=n0, n1=n1: self._encode_head(sx[n0:n1], y1.batch_slice(n0, n1)),
                                             cache=False)
                    heads.append(g)
                    launches += l

                def tail():
                    sh.zero_()
                    return self._encode_tail(y1, sh)

                gt, _, codes, l = self._graph(key + ("tail",), list, tail, cache=False)
                pipe = dict(sx=sx, sh=sh, y1=y1, heads=heads, tail=gt, codes=codes, launches=launches + l,
                            events=[torch.cuda.Event() for _ in bounds], done=torch.cuda.Event())
                self._pipes.put(key, pipe, torch.cuda.memory_allocated() - mem0)
            main, copy = self._streams(dev)
            copy.wait_stream(main)          # the previous step's graphs may still read the staging buffer
            with torch.cuda.stream(copy):
                for c, (n0, n1) in enumerate(bounds):
                    pipe["sx"][n0:n1].copy_(x[n0:n1], non_blocking=True)
                    pipe["events"][c].record(copy)
                pipe["done"].record(copy)
            for c in range(len(bounds)):
                main.wait_event(pipe["events"][c])
                pipe["heads"][c].replay()
            pipe["tail"].replay()
            self.graph_launches += pipe["launches"]
            if hist is not None:
                hist += pipe["sh"]
            out = [c.clone() for c in pipe["codes"]]
            # the caller may refill its pinned batch as soon as encode() returns (double buffering): all H2D copies of
            # `x` have completed by then (they finished long before the graphs queued behind them do)
            pipe["done"].synchronize()
            return out

    def _decode_pipelined(self, codes: List[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
        """out: pinned host batch.  The pixels of chunk c travel to the host while chunk c+1 runs the last layers."""
        dev = codes[0].device
        n = codes[0].shape[0]
        bounds = self.host_slices(n, small_first=False, itemsize=out.element_size())
        key = ("dec", tuple(tuple(c.shape) for c in codes), tuple(out.shape), out.dtype, self.decode_passes, dev)
        pipe = self._pipes.get(key)
        with torch.cuda.device(dev):
            if pipe is None:
                mem0 = torch.cuda.memory_allocated()
                sc = [torch.zeros_like(c) for c in codes]
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                sout = torch.empty(tuple(out.shape), dtype=out.dtype, device=dev)

                def main_body():
                    status.zero_()
                    return self._decode_main(sc, status)

                gm, _, y4, launches = self._graph(key + ("main",), list, main_body, cache=False)
                if (y4.n, 2 * y4.h, 2 * y4.w) != (out.shape[0], out.shape[2], out.shape[3]) or out.shape[1] != 3:
                    raise RuntimeError(f"`out` must be [n, 3, H_pad, W_pad] = [{y4.n}, 3, {2 * y4.h}, {2 * y4.w}]")
                tails = []
                for c, (n0, n1) in enumerate(bounds):
                    g, _, _, l = self._graph(key + ("tail", c), list,
                                             lambda n0=n0, n1=n1: self._decode_tail(y4.batch_slice(n0, n1), sout[n0:n1]),
                                             cache=False)
                    tails.append(g)
                    launches += l
                pipe = dict(sc=sc, status=status, sout=sout, main=gm, tails=tails, y4=y4, launches=launches,
                            events=[torch.cuda.Event() for _ in bounds])
                self._pipes.put(key, pipe, torch.cuda.memory_allocated() - mem0)
            main, copy = self._streams(dev)
            for dst, src in zip(pipe["sc"], codes):
                dst.copy_(src)
            pipe["main"].replay()
            for c, (n0, n1) in enumerate(bounds):
                pipe["tails"][c].replay()
                pipe["events"][c].record(main)
                with torch.cuda.stream(copy):
                    copy.wait_event(pipe["events"][c])
                    out[n0:n1].copy_(pipe["sout"][n0:n1], non_blocking=True)
            main.wait_stream(copy)
            self.graph_launches += pipe["launches"]
            if int(pipe["status"].item()) != 0:      # also the synchronisation point: `out` is complete on return
                raise RuntimeError("code index out of range for its codebook")
        return out

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def encode(self, x: torch.Tensor, hist: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """compressor.py:79-88.  `hist`: optional flat int32 [sum_l m*k_l] code histogram, accumulated in place.
        x may also be a pinned host batch: it is then streamed to the GPU in chunks that overlap the first layers
        (codes are returned on the model's device), and / or uint8 RGB images (host or device) as `demo.compressImage`
        receives them (demo.py:109-118): `convert_image_dtype` + `(x - 0.5) * 2` then happen inside the first kernel
        with the reference's fp32 operations, so a host batch crosses PCIe as bytes."""
        self._check_image(x)
        self._check_weights()
        if not x.is_cuda and self._host_batch_ok(x):
            return self._encode_pipelined(x, hist)
        if not (self.use_graphs and x.is_cuda):
            return self._encode_eager(x, hist)
        x = x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()
        total = self._quantizer.hist_size()

        def make_static():
            return [torch.empty_like(x), torch.zeros(total, dtype=torch.int32, device=x.device)]

        def body(sx, sh):
            sh.zero_()
            return self._encode_eager(sx, sh)

        graph, (sx, sh), codes, launches = self._graph(("enc", tuple(x.shape), x.dtype, self.encode_passes, x.device,
                                                        self.encode_slice), make_static, body)
        sx.copy_(x)
        graph.replay()
        self.graph_launches += launches
        if hist is not None:
            hist += sh
        return [c.clone() for c in codes]

    @torch.no_grad()
    def decode(self, codes: List[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """compressor.py:114-117 (no crop; `decompress` crops upstream).
        out (extension): a pinned host tensor [n, 3, H_pad, W_pad]; the pixels are then streamed into it in chunks
        that overlap the last layers, and `out` is returned (complete when the call returns).  fp32: the pixels as upstream;
        uint8: the pixels through the reference's DeTransform (utils/vision.py:135-146, what demo.decompressImage returns),
        applied in the last kernel's epilogue."""
        if len(codes) == 0:
            raise RuntimeError("Length of codes is 0.")
        self._check_weights()
        dev = codes[0].device
        if out is not None:
            ok = all(c.is_cuda and c.dtype == torch.int64 and c.dim() == 4 and c.is_contiguous() for c in codes)
            if not (ok and self._host_batch_ok(out) and out.shape[0] == codes[0].shape[0]):
                raise RuntimeError("decode(out=): needs CUDA int64 codes and a pinned contiguous fp32 / uint8 host tensor "
                                   "[n, 3, H_pad, W_pad]")
            if len(codes) != len(self._quantizer._k):
                raise RuntimeError(f"expected {len(self._quantizer._k)} code levels, got {len(codes)}")
            return self._decode_pipelined(codes, out)
        if not (self.use_graphs and codes[0].is_cuda):
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            out = self._decode_eager(codes, status)
        else:
            codes = [c.contiguous() for c in codes]
            for c in codes:
                if c.dtype != torch.int64 or c.dim() != 4:
                    raise RuntimeError(f"codes must be int64 [n, m, h, w], got {c.dtype} {tuple(c.shape)}")

            def make_static():
                return [[torch.empty_like(c) for c in codes], torch.zeros(1, dtype=torch.int32, device=dev)]

            def body(sc, st):
                st.zero_()
                return self._decode_eager(sc, st)

            key = ("dec", tuple(tuple(c.shape) for c in codes), self.decode_passes, dev, self.decode_slice)
            graph, (sc, status), sout, launches = self._graph(key, make_static, body)
            for dst, src in zip(sc, codes):
                dst.copy_(src)
            graph.replay()
            self.graph_launches += launches
            out = sout.clone()
        if int(status.item()) != 0:
            raise RuntimeError("code index out of range for its codebook")
        return out

    @torch.no_grad()
    def compress(self, x: torch.Tensor):
        """compressor.py:67-77: (codes, binaries [n][L] bytes, n FileHeader records).  encode on the GPU, rANS on the
        host (one batched multi-threaded call per level)."""
        from .. import entropy
        from ..container import REFERENCE_VERSION as __version__   # header version = the reference release we interoperate with
        n, c, h, w = x.shape
        codes = self.encode(x)
        binaries, sizes = self._quantizer._entropyCoder.compress(codes)
        headers = [entropy.FileHeader(__version__, self._qp, size, entropy.ImageSize(height=h, width=w, channel=c))
                   for size in sizes]
        return codes, binaries, headers

    @torch.no_grad()
    def decompress(self, binaries, headers) -> torch.Tensor:
        """compressor.py:90-112: rANS decode on the host, synthesis on the GPU, centre-crop to the header's image size."""
        codes = self._quantizer._entropyCoder.decompress(binaries, [hd.CodeSize for hd in headers])
        restored = self.decode(codes)
        size = headers[0].ImageSize
        H, W = restored.shape[-2], restored.shape[-1]
        if not (0 < size.height <= H and 0 < size.width <= W):      # the header comes from a file
            raise RuntimeError(f"decompress: header image size {size.height} x {size.width} does not fit the {H} x {W} "
                               "pixels its code maps decode to")
        top, left = (H - size.height) // 2, (W - size.width) // 2
        return restored[..., top:top + size.height, left:left + size.width]

    @property
    def CDFs(self):
        return self._quantizer.CDFs

    def _analysis_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """y = self._encoder(x) on the engine, NCHW fp32 in / out (no padding: compressor.py:35-43 takes training crops)"""
        eng = self.engine
        n, _, h, w = x.shape
        y0 = eng.stem(self._encoder[0], x, (0, 0, h, w), eng.needs_of(self._encoder[1]))
        return eng.to_nchw(eng.run_seq(list(self._encoder)[1:], y0, {"f32"}))

    def _synthesis_nchw(self, yHat: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        act = eng.from_nchw(yHat, eng.needs_of(self._decoder[0]))
        return eng.run_seq(list(self._decoder), act, set()).f32        # the final pixel-shuffle conv stores NCHW

    def forward(self, x: torch.Tensor):
        """compressor.py:35-43: (xHat, yHat, codes, logits) of the training-time path (soft quantizer: Gumbel sample with
        PyTorch's RNG, frequency EMA updated).
        * module in training mode and autograd enabled: the TRAINING STEP (SURVEY 8f NEXT-3) -- an autograd graph is built,
          every convolution runs forward / dgrad / wgrad on the tcgen05 kernels (mcquic_b200/autograd.py);
          `loss(xHat, x).backward()` fills the parameters' `.grad` like upstream.
        * otherwise (eval mode or torch.no_grad()): FORWARD VALUES ONLY on the inference engine, for monitoring /
          validation of a training run.  (Upstream returns nothing in eval mode; here the values are returned.)"""
        self._check_image(x)
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise RuntimeError("forward() takes training crops whose sides the strided stages divide (multiples of 16)")
        if self.training and torch.is_grad_enabled():
            from ..autograd import compressor_forward
            return compressor_forward(self, x)
        with torch.no_grad():
            return self._forward_values(x)

    def _forward_values(self, x: torch.Tensor):
        from .. import engine as E
        old, E._DEFAULT = E._DEFAULT, self.engine       # the quantizer's values path runs on this model's engine
        try:
            self.engine.passes = self.encode_passes
            y = self._analysis_nchw(x.contiguous().float())
            yHat, codes, logits = self._quantizer(y)
            xHat = self._synthesis_nchw(yHat)
            self.engine.flush()
        finally:
            E._DEFAULT = old
        return xHat, yHat, codes, logits


class Compressor(BaseCompressor):
    def __init__(self, channel: int, m: int, k: List[int], permutationRate: float = 0.0):
        if channel % 8 != 0:
            raise ValueError("channel must be a multiple of 8")
        RB, RBS, RBU, AB = ResidualBlock, ResidualBlockWithStride, ResidualBlockShuffle, AttentionBlock
        C = channel
        encoder = nn.Sequential(conv3x3(3, C, 2), RB(C, C), RBS(C, C), AB(C), RB(C, C), RBS(C, C), RB(C, C))
        decoder = nn.Sequential(RB(C, C), RBU(C, C), AB(C), RB(C, C), RBU(C, C), RB(C, C), pixelShuffle3x3(C, 3, 2))
        quantizer = UMGMQuantizer(C, m, k, permutationRate, {
            "latentStageEncoder": lambda: nn.Sequential(RBS(C, C), RB(C, C), AB(C)),
            "quantizationHead": lambda: nn.Sequential(RB(C, C), AB(C), conv3x3(C, C)),
            "latentHead": lambda: nn.Sequential(RB(C, C), AB(C), conv3x3(C, C)),
            "restoreHead": lambda: nn.Sequential(AB(C), RB(C, C), RBU(C, C)),
            "dequantizationHead": lambda: nn.Sequential(AB(C), conv3x3(C, C), RB(C, C)),
            "sideHead": lambda: nn.Sequential(AB(C), conv3x3(C, C), RB(C, C)),
        })
        super().__init__(encoder, quantizer, decoder)


class Neon(BaseCompressor):
    """The tokenizer of BASELINE configs[4] (mcquic/modules/compressor.py:181-233): a full-resolution stem (no stride),
    three strided stages, a bottleneck to the quantizer's 8 channels, and `ResidualBackwardQuantizer` (len(size) levels
    sharing one [1, k, 8] codebook).  With denseNorm=True every ResidualBlock normalises with nn.GroupNorm (32 groups in
    the trunk, 1 around the quantizer).  encode/decode (inference) run on the CUDA engine; `codes` are ordered smallest
    level first, as upstream."""

    def __init__(self, channel: int, k: int, size: List[int], denseNorm: bool = False, *_, **__):
        quantizer = ResidualBackwardQuantizer(k, list(size), denseNorm)
        RB, RBS, RBU, AB = ResidualBlock, ResidualBlockWithStride, ResidualBlockShuffle, AttentionBlock
        C, Q, d = channel, quantizer.channel, denseNorm
        encoder = nn.Sequential(
            conv3x3(3, C), AB(C, 32, d), RB(C, C, 32, d), RB(C, C, 32, d), RBS(C, C, 2, 32, d), RB(C, C, 32, d),
            RBS(C, C, 2, 32, d), RB(C, C, 32, d), RBS(C, C, 2, 32, d), AB(C, 32, d), RB(C, 2 * C, 32, d),
            RB(2 * C, 2 * C, 32, d), RB(2 * C, 2 * C, 32, d), RB(2 * C, 2 * C, 32, d), RB(2 * C, Q, 1, d), AB(Q, 1, d))
        decoder = nn.Sequential(
            AB(Q, 1, d), RB(Q, 2 * C, 1, d), RB(2 * C, 2 * C, 32, d), RB(2 * C, 2 * C, 32, d), RB(2 * C, 2 * C, 32, d),
            RB(2 * C, C, 32, d), AB(C, 32, d), RB(C, C, 32, d), RBU(C, C, 2, 32, d), RB(C, C, 32, d), RBU(C, C, 2, 32, d),
            RB(C, C, 32, d), RBU(C, C, 2, 32, d), RB(C, C, 32, d), RB(C, C, 32, d), AB(C, 32, d), conv3x3(C, 3))
        super().__init__(encoder, quantizer, decoder)
        # 57 convolutions and 50 GroupNorms deep: the single-pass (TF32-grade) decode drifts to ~2.5e-3 of the output
        # range, so the tokenizer decodes with the fp32-grade 3-pass path unless told otherwise
        self.decode_passes = 3

    def _host_batch_ok(self, t: torch.Tensor) -> bool:
        return False        # the host I/O pipeline is built around Compressor's strided stem / pixel-shuffle tail

    def _encode_eager(self, x: torch.Tensor, hist: Optional[torch.Tensor]) -> List[torch.Tensor]:
        if x.dtype == torch.uint8:
            raise RuntimeError("Neon.encode takes float images in [-1, 1] (uint8 input is a Compressor extension)")
        eng = self.engine
        eng.passes = self.encode_passes
        n, _, h, w = x.shape
        top, left, hp, wp = aligned_pad_amounts(h, w)
        if (hp, wp) != (h, w):       # AlignedPadding (transforms.py:86-99): a copy with reflected borders, no arithmetic
            x = torch.nn.functional.pad(x, (left, wp - w - left, top, hp - h - top), "reflect")
        a0 = eng.from_nchw(x, eng.needs_of(self._encoder[0]), pad_channels_to=8)
        y = eng.run_seq(list(self._encoder), a0, self._quantizer.first_needs(eng))
        codes = self._quantizer.encode_act(eng, y, hist)
        eng.flush()
        return codes

    def _decode_eager(self, codes: List[torch.Tensor], status: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        eng.passes = self.decode_passes
        yHat = self._quantizer.decode_act(eng, codes, eng.needs_of(self._decoder[0]), status)
        out = eng.run_seq(list(self._decoder), yHat, {"f32"})      # final conv C -> 3 runs with 8 stored channels
        eng.flush()
        return eng.to_nchw(out)[:, :3].contiguous()

    def _analysis_nchw(self, x: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        a0 = eng.from_nchw(x, eng.needs_of(self._encoder[0]), pad_channels_to=8)
        return eng.to_nchw(eng.run_seq(list(self._encoder), a0, {"f32"}))

    def _synthesis_nchw(self, yHat: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        act = eng.from_nchw(yHat, eng.needs_of(self._decoder[0]))
        return eng.to_nchw(eng.run_seq(list(self._decoder), act, {"f32"}))[:, :3].contiguous()

    def residual_backward(self, code: torch.Tensor, level: int) -> torch.Tensor:
        return self._quantizer.residual_backward(code, level)       # compressor.py:235-237

    def residual_forward(self, code: torch.Tensor, formerLevel: Optional[torch.Tensor], level: int) -> torch.Tensor:
        return self._quantizer.residual_forward(code, formerLevel, level)   # compressor.py:239-241

    def compress(self, x: torch.Tensor):
        raise NotImplementedError("VariousMCoder.compress raises upstream as well (entropyCoder.py:250)")

    def decompress(self, binaries, headers):
        raise NotImplementedError("VariousMCoder.decompress raises upstream as well (entropyCoder.py:281)")
