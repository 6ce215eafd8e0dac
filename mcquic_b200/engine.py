"""Host-side executor: walks the reference-shaped module tree and issues one fused CUDA launch per convolution.

PyTorch is used for device memory, streams and (one-time) weight repacking only; every FLOP of the hot path
runs in libmcquic_b200.so (`_lib.py`).  Fusion map (reference op -> where it went):

  nn.SiLU before a conv              -> producer epilogue writes silu(y) as a split-fp16 plane pair
  `out += identity` (blocks.py:77)   -> residual operand of the consuming conv's epilogue
  GDN / IGDN (gdn.py:67-91)          -> 1x1 tensor-core conv over x^2 planes, epilogue x * rsqrt/sqrt(.)
  PixelShuffle (convs.py:252-255)    -> store addressing of the producing conv (weights row-permuted at load)
  a * sigmoid(b) + x (blocks.py:281) -> epilogue of the AttentionBlock's trailing 1x1 conv
  nn.GroupNorm (blocks.py:198)       -> statistics: (sum, sum^2) partials written by the producing conv's epilogue
                                        (conv_pair_kernel<.., GN>); normalise + affine + split into the next conv's
                                        operand planes: one streaming pass (mcq_groupnorm_apply).  Shapes the fused path
                                        does not take: one cluster-per-image launch doing all of it (mcq_groupnorm)
  conv1x1 skip (blocks.py:189-192)   -> its fp32 output is the residual operand of the block's second conv
  z - dequant(code) (quantizer.py:318), q + side (quantizer.py:354) -> residual operands
  AlignedPadding (transforms.py:86)  -> index arithmetic of the stem kernel
"""
import contextlib
import ctypes
import weakref
import os
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Set, Tuple

import torch
from torch import nn

from . import _lib
from .nn.blocks import AttentionBlock, ResidualBlock, ResidualBlockShuffle, ResidualBlockWithStride
from .nn.gdn import GenDivNorm

Planes = Tuple[torch.Tensor, Optional[torch.Tensor]]  # (hi, lo) fp16 NHWC; lo is None on the 1-pass path

LO_SCALE = 2048.0


@dataclass
class Act:
    """One activation, NHWC, in whichever representations its consumers need."""
    n: int
    h: int
    w: int
    c: int
    f32: Optional[torch.Tensor] = None
    raw: Optional[Planes] = None
    silu: Optional[Planes] = None
    sq: Optional[Planes] = None
    # (partials float2 [n, rowblocks, units], rowblocks, unit): GroupNorm statistics the producing conv's epilogue wrote
    gn: Optional[Tuple[torch.Tensor, int, int]] = None

    def batch_slice(self, n0: int, n1: int) -> "Act":
        """Images [n0, n1) of this activation (NHWC: a contiguous view of every representation)."""
        def cut(t):
            return None if t is None else t[n0:n1]
        def cut2(pl):
            return None if pl is None else (cut(pl[0]), cut(pl[1]))
        return Act(n1 - n0, self.h, self.w, self.c, cut(self.f32), cut2(self.raw), cut2(self.silu), cut2(self.sq))


@dataclass
class PackedConv:
    w_hi: torch.Tensor
    w_lo: torch.Tensor
    bias: torch.Tensor
    cin: int
    cout: int
    cout_pad: int
    ksize: int
    stride: int
    w_scale: float
    store: int


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def split_weight(w2d: torch.Tensor, amax: Optional[float] = None, exp: Optional[int] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """fp32 [rows, K] -> (hi, lo) fp16 with  w * 2^e ~= hi + lo / 2048  and the accumulator scale 2^-e.
    amax: max |w| when the caller already knows it (Engine.prepare computes it for ALL layers with one device
    reduction and one host read instead of a device-to-host sync per layer).
    exp: the exponent e itself (training step: chosen when a layer is first packed and kept while its weights drift, so
    that re-packing after every optimizer step needs no host synchronisation and can be captured in a CUDA graph).
    The arithmetic is exact in fp32 (power-of-two scaling; w * 2^e - hi has at most 13 significant bits)."""
    if exp is None:
        if amax is None:
            amax = float(w2d.abs().max())
        exp = 0 if amax == 0.0 or not math.isfinite(amax) else int(max(-14, min(14, math.floor(math.log2(256.0 / amax)))))
    ws = w2d.float() * (2.0 ** exp)
    hi = ws.to(torch.float16)
    lo = ((ws - hi.float()) * LO_SCALE).to(torch.float16)
    return hi.contiguous(), lo.contiguous(), float(2.0 ** (-exp))


def pack_codebook(codebook: torch.Tensor):
    """[m, k, d] fp32 -> (hi, lo, scale, lohi): split-fp16 planes [m*k, d] of c * 2^e and the fused-VQ operand
    [m*k, 2d] whose rows are [lo | hi] (csrc/vq_fused.cuh)."""
    d = codebook.shape[-1]
    hi, lo, scale = split_weight(codebook.detach().float().reshape(-1, d))
    return hi, lo, scale, torch.cat([lo, hi], dim=1).contiguous()


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, store: int, device,
              amax: Optional[float] = None, exp: Optional[int] = None) -> PackedConv:
    """nn.Conv2d weight [cout, cin, k, k] -> K-major GEMM matrix [cout_pad, (r, s, cin)] (split fp16).
    bias=None (conv1x1(..., bias=False) of the Neon quantizer) packs a zero bias.  Channel counts the kernels' vector
    accesses cannot address are zero-padded: cin to a multiple of 8 (RGB input of Neon's first conv: the caller pads
    the activation likewise; 8 fp16 channels = the 16 B row alignment TMA needs, so the layer runs on the tensor cores)
    and, for plain NHWC stores, cout to a multiple of 8 (Neon's final C -> 3 conv: the
    caller drops the extra channels); `PackedConv.cin / .cout` are the padded counts."""
    cout, cin, k, _ = weight.shape
    w4 = weight.detach().to(device=device, dtype=torch.float32)
    b = torch.zeros(cout, device=device) if bias is None else bias.detach().to(device=device, dtype=torch.float32)
    if cin % 4 != 0:
        w4 = torch.cat([w4, torch.zeros(cout, 8 - cin % 8, k, k, device=device)], 1)
        cin = w4.shape[1]
    if store == _lib.STORE_NHWC and cout % 8 != 0:
        extra = 8 - cout % 8
        w4 = torch.cat([w4, torch.zeros(extra, cin, k, k, device=device)], 0)
        b = torch.cat([b, torch.zeros(extra, device=device)])
        cout += extra
    w = w4.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
    if store == _lib.STORE_SHUFFLE_NHWC:
        # PixelShuffle(2): conv channel 4c + 2i + j -> GEMM column (2i + j) * C + c, so an N tile is one sub-pixel
        cq = cout // 4
        perm = torch.arange(cout, device=device).reshape(cq, 4).t().reshape(-1)
        w, b = w[perm], b[perm]
    cout_pad = cout if cout % 128 == 0 or (cout <= 256 and cout % 32 == 0) else ((cout + 15) // 16) * 16
    if cout_pad != cout:
        w = torch.cat([w, torch.zeros(cout_pad - cout, w.shape[1], device=device)], 0)
    hi, lo, scale = split_weight(w, amax, exp)   # zero padding and row permutation leave max |w| unchanged
    return PackedConv(hi, lo, b.contiguous(), cin, cout, cout_pad, k, stride, scale, store)


class Engine:
    """passes: 3 = split-fp16 fp32-grade (encode), 1 = single fp16 pass, TF32-grade (decode).
    impl: 'tcgen05' (tensor cores) or 'simt' (fp32 CUDA-core cross-check kernel)."""

    def __init__(self, impl: str = "tcgen05", lib=None):
        # `lib` is a test seam (tests/emulator.py injects a CPU model of the C ABI to check this host logic
        # without a GPU); the product always binds the CUDA library and raises if it cannot be loaded.
        self.emulated = lib is not None
        self.lib = lib if lib is not None else _lib.load()
        self.impl = {"tcgen05": _lib.IMPL_TCGEN05, "simt": _lib.IMPL_SIMT}[impl]
        self.passes = 3
        self._packed: Dict[Tuple[int, int], Tuple] = {}   # (id(module), store) -> (weight version, packing, weakref(module))
        # independent sub-graphs (the two AttentionBlock branches, quantizationHead || latentHead,
        # dequantizationHead || sideHead) run on side streams: on <=16x16 feature maps one conv cannot fill 148 SMs
        self.multistream = not self.emulated
        self._side_streams: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}
        self._depth = 0
        # when a list, every conv launch is bracketed by CUDA events and logged (bench.py's roofline leg)
        self.profile: Optional[list] = None
        # chain=True: convolutions on <= 16x16 maps are not launched one by one but collected and issued as one
        # persistent layer-chain launch (mcq_conv_chain).  Measured on B200 a chained layer costs
        # as much as a stand-alone launch that overlaps a second stream, so the default stays layer-by-layer launches
        # on two streams; the emulated ABI keeps chains on so that the CPU tests cover the recording/merge logic.
        self.chain = self.emulated
        self._pending: list = []
        self._rec_depth = 0
        # stride-1 convs with cin % 8 == 0 run on the tensor cores with a partly zero-filled last K chunk (A/B knob)
        self.partial_chunk = True
        self.stem_tc = True      # the stem on the tensor cores (csrc/stem_tc.cuh); False: FFMA kernel (A/B knob)

    # ------------------------------------------------------------------ plumbing
    @contextlib.contextmanager
    def _prof(self, name: str):
        """When profiling, bracket a non-convolution launch (stem, VQ, gather, GroupNorm, layout, ...) by events too:
        entries {"other": name, "ev": (start, stop)} -- bench.py's share of the step is taken over ALL launches."""
        if self.profile is None:
            yield
            return
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        yield
        ev1.record()
        self.profile.append({"other": name, "ev": (ev0, ev1)})

    def _stream(self):
        if self.emulated:
            return ctypes.c_void_p(0)
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def can_chain(self, x: "Act") -> bool:
        """True if every conv of a block fed with x goes into the pending layer chain."""
        return (self.chain and self.impl == _lib.IMPL_TCGEN05 and x.h * x.w <= self.CHAIN_MAX_PIXELS
                and x.c % 64 == 0)

    def parallel(self, fa, fb, chain: bool = False):
        """Run the independent closures fa and fb concurrently; returns (fa(), fb()).
        chain=True (caller guarantees both consist of chainable convs only): their layers are interleaved in the
        pending chain, so the kernel runs them without a barrier in between.  Otherwise fb goes to a side stream."""
        if chain:
            outer = self._pending
            self._rec_depth += 1
            try:
                self._pending = la = []
                ra = fa()
                self._pending = lb = []
                rb = fb()
            finally:
                self._pending = outer
                self._rec_depth -= 1
            for i in range(max(len(la), len(lb))):
                outer.extend(la[i:i + 1])
                outer.extend(lb[i:i + 1])
            return ra, rb
        if not self.multistream:
            return fa(), fb()
        # one side stream per (nesting depth, parent stream): a side stream shared by two parents would let the
        # caching allocator hand a block to the second parent's branch while the first parent still reads it
        self.flush()
        main = torch.cuda.current_stream()
        key = (self._depth, main.cuda_stream)
        side = self._side_streams.get(key)
        if side is None:
            side = self._side_streams[key] = torch.cuda.Stream()
        self._depth += 1
        try:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rb = fb()
                self.flush()
            ra = fa()
            self.flush()
            main.wait_stream(side)
        finally:
            self._depth -= 1
        return ra, rb

    # ------------------------------------------------------------------ pending layer chain
    CHAIN_MAX_PIXELS = 256   # conv output grids up to 16x16 go into chains

    def _launch_one(self, item):
        p, _keep, info = item
        if self.profile is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            ev1.record()   # forces creation of the underlying cudaEvent_t handles
            p.ev_start, p.ev_stop = ev0.cuda_event, ev1.cuda_event
        rc = self.lib.mcq_conv2d(ctypes.byref(p), self._stream())
        if rc == _lib.ERR_UNSUPPORTED and p.gn_partials:
            # the statistics-fusing instantiation cannot take this layer after all (shared-memory budget): plain launch,
            # the GroupNorm that follows computes its own statistics
            p.gn_partials = None
            _keep[1].gn = None
            rc = self.lib.mcq_conv2d(ctypes.byref(p), self._stream())
        _lib.check(rc, "mcq_conv2d")
        if self.profile is not None:
            self.profile.append(dict(info, ev=(ev0, ev1)))

    def _launch_group(self, items):
        if len(items) == 1:
            return self._launch_one(items[0])
        arr = (_lib.ConvParams * len(items))(*[it[0] for it in items])
        if self.profile is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            ev1.record()
            arr[0].ev_start, arr[len(items) - 1].ev_stop = ev0.cuda_event, ev1.cuda_event
        rc = self.lib.mcq_conv_chain(arr, len(items), self._stream())
        if rc == _lib.ERR_UNSUPPORTED:      # nothing was launched: split (too many layers / bias pool) or go one by one
            half = len(items) // 2
            self._launch_group(items[:half])
            self._launch_group(items[half:])
            return
        _lib.check(rc, "mcq_conv_chain")
        if self.profile is not None:
            self.profile.append({"flops": sum(it[2]["flops"] for it in items), "passes": items[0][2]["passes"],
                                 "impl": _lib.IMPL_TCGEN05, "ev": (ev0, ev1),
                                 "shape": ("chain", len(items)) + tuple(items[0][2]["shape"][1:3]),
                                 "layers": [it[2]["shape"] for it in items]})

    def flush(self):
        """Issue the pending layer chain on the current stream."""
        if self._rec_depth:
            raise RuntimeError("mcquic_b200: non-chainable op inside a chain-recorded parallel region")
        items, self._pending = self._pending, []
        if not items:
            return
        maxl = int(self.lib.mcq_conv_chain_max_layers())
        for i in range(0, len(items), maxl):
            self._launch_group(items[i:i + maxl])

    def _planes(self, n, h, w, c, device) -> Planes:
        hi = torch.empty((n, h, w, c), dtype=torch.float16, device=device)
        lo = torch.empty((n, h, w, c), dtype=torch.float16, device=device) if self.passes == 3 else None
        return hi, lo

    def alloc_act(self, n, h, w, c, want: Set[str], device) -> Act:
        """Caller-owned output buffers for `conv(..., into=)` (chunked execution writes batch slices of them)."""
        out = Act(n, h, w, c)
        if "f32" in want:
            out.f32 = torch.empty((n, h, w, c), dtype=torch.float32, device=device)
        for name in ("raw", "silu", "sq"):
            if name in want:
                setattr(out, name, self._planes(n, h, w, c, device))
        return out

    def _packed_for(self, mod: nn.Module, store: int = _lib.STORE_NHWC, amax: Optional[float] = None) -> PackedConv:
        key = (id(mod), store)
        def version(t):          # tensors created under torch.inference_mode() (the reference CLI runs that way,
            try:                 # mcquic/cli.py:60) have no version counter: they are immutable, so 0 is exact
                return t._version
            except RuntimeError:
                return 0

        if isinstance(mod, GenDivNorm):
            ver = (version(mod.beta), version(mod.gamma), mod.beta.data_ptr(), mod.gamma.data_ptr(), tuple(mod.gamma.shape))
        elif mod.bias is None:
            ver = (version(mod.weight), mod.weight.data_ptr(), tuple(mod.weight.shape))
        else:
            ver = (version(mod.weight), version(mod.bias), mod.weight.data_ptr(), mod.bias.data_ptr(),
                   tuple(mod.weight.shape))
        hit = self._packed.get(key)
        # (id(mod) can come back with another module once `mod` is collected -- and with it, through the caching allocator,
        #  the weight's address and version: the entry also remembers which module it was packed for)
        if hit is not None and hit[0] == ver and hit[2]() is mod:
            return hit[1]
        if isinstance(mod, GenDivNorm):
            beta, gamma = mod.effective()
            pc = pack_conv(gamma[:, :, None, None], beta, 1, _lib.STORE_NHWC, gamma.device, amax)
        else:
            pc = pack_conv(mod.weight, mod.bias, mod.stride[0], store, mod.weight.device, amax)
        self._packed[key] = (ver, pc, weakref.ref(mod))
        return pc

    def prepare(self, root: nn.Module):
        """Repack every convolution / GDN under `root` now (otherwise done lazily on first use, with one host sync per
        layer for its max |w|): max |w| of ALL layers comes from one batched device reduction and one host read."""
        jobs = []        # (module, store)
        shuffled = set()
        for mod in root.modules():
            if isinstance(mod, ResidualBlockShuffle):
                shuffled.update((id(mod._branch[1][0]), id(mod._skip[0])))
        final = None
        for mod in root.modules():
            if isinstance(mod, nn.Sequential) and len(mod) == 2 and isinstance(mod[1], nn.PixelShuffle) \
                    and isinstance(mod[0], nn.Conv2d) and id(mod[0]) not in shuffled:
                final = id(mod[0])      # stand-alone pixelShuffle3x3 = last decoder layer (NCHW pixel store)
        for mod in root.modules():
            if isinstance(mod, GenDivNorm):
                jobs.append((mod, _lib.STORE_NHWC))
            elif isinstance(mod, nn.Conv2d) and mod.kernel_size[0] in (1, 3) and mod.in_channels != 3:
                store = _lib.STORE_SHUFFLE_NHWC if id(mod) in shuffled else (
                    _lib.STORE_SHUFFLE_NCHW if id(mod) == final else _lib.STORE_NHWC)
                jobs.append((mod, store))
        if not jobs:
            return
        with torch.no_grad():
            ws = [(m.effective()[1] if isinstance(m, GenDivNorm) else m.weight).detach().float() for m, _ in jobs]
            amax = torch.stack(torch._foreach_norm(ws, float("inf"))).tolist()       # the ONE host sync
        for (mod, store), am in zip(jobs, amax):
            self._packed_for(mod, store, amax=float(am))

    # ------------------------------------------------------------------ one fused conv launch
    def conv(self, pc: PackedConv, a: Planes, x: Act, want: Set[str], *, mode: int = _lib.EPI_LINEAR,
             res1: Optional[torch.Tensor] = None, res1_scale: float = 1.0, res2: Optional[torch.Tensor] = None,
             aux: Optional[torch.Tensor] = None, into: Optional[Act] = None, gn_groups: int = 0,
             a_scale: float = 1.0, dev_scale: Optional[torch.Tensor] = None) -> Act:
        """a_scale: power-of-two factor the A planes carry (x^2 planes: _lib.SQUARE_SCALE); undone through w_scale.
        dev_scale: optional device scalar (fp32 [1]) the epilogue multiplies into w_scale (training-step dgrad).
        gn_groups > 0: a nn.GroupNorm(gn_groups, cout) follows; where the kernel can (mcq_conv_gn_layout) its
        epilogue also writes the per-row-block (sum, sum^2) partials of the fp32 output -> `out.gn`.
        into: write the outputs into these caller-owned tensors (same shapes / representations as `want`) instead
        of allocating them -- used when a batch is processed in chunks that fill slices of one full-batch tensor."""
        assert pc.cin == x.c, (pc.cin, x.c)
        dev = a[0].device
        ho, wo = x.h // pc.stride, x.w // pc.stride
        co = pc.cout
        if pc.store != _lib.STORE_NHWC:
            ho, wo, co = ho * 2, wo * 2, pc.cout // 4
        out = Act(x.n, ho, wo, co)
        p = _lib.ConvParams()
        p.a_hi, p.a_lo = _ptr(a[0]), _ptr(a[1])
        p.n, p.hin, p.win, p.cin = x.n, x.h, x.w, x.c
        p.w_hi, p.w_lo = _ptr(pc.w_hi), _ptr(pc.w_lo)
        p.cout, p.cout_pad, p.ksize, p.stride = pc.cout, pc.cout_pad, pc.ksize, pc.stride
        p.w_scale = pc.w_scale / a_scale
        p.dev_scale = _ptr(dev_scale)
        p.bias = _ptr(pc.bias)
        p.mode, p.store = mode, pc.store
        p.res1, p.res1_scale, p.res2, p.aux = _ptr(res1), res1_scale, _ptr(res2), _ptr(aux)
        keep = [a, res1, res2, aux, dev_scale]
        def given(t, shape, dtype):
            if t is None or tuple(t.shape) != shape or t.dtype != dtype or not t.is_contiguous():
                raise RuntimeError("mcquic_b200: `into` buffers do not match the convolution's outputs")
            return t

        if pc.store == _lib.STORE_SHUFFLE_NCHW:
            if into is not None and into.f32 is not None and into.f32.dtype == torch.uint8:
                # uint8 pixels through the reference's DeTransform (vision.py:135-146) in the epilogue
                out.f32 = given(into.f32, (x.n, co, ho, wo), torch.uint8)
                p.out_u8 = _ptr(out.f32)
            else:
                if into is not None:
                    out.f32 = given(into.f32, (x.n, co, ho, wo), torch.float32)
                else:
                    out.f32 = torch.empty((x.n, co, ho, wo), dtype=torch.float32, device=dev)  # NCHW pixels
                p.out_f32 = _ptr(out.f32)
        else:
            if "f32" in want:
                if into is not None:
                    out.f32 = given(into.f32, (x.n, ho, wo, co), torch.float32)
                else:
                    out.f32 = torch.empty((x.n, ho, wo, co), dtype=torch.float32, device=dev)
                p.out_f32 = _ptr(out.f32)
            slots = []
            for name, act in (("raw", _lib.ACT_NONE), ("silu", _lib.ACT_SILU), ("sq", _lib.ACT_SQUARE)):
                if name in want:
                    if into is not None:
                        pl = getattr(into, name)
                        if pl is None:
                            raise RuntimeError(f"mcquic_b200: `into` lacks the '{name}' planes")
                        given(pl[0], (x.n, ho, wo, co), torch.float16)
                        if self.passes == 3:
                            given(pl[1], (x.n, ho, wo, co), torch.float16)
                        else:
                            pl = (pl[0], None)
                    else:
                        pl = self._planes(x.n, ho, wo, co, dev)
                    setattr(out, name, pl)
                    slots.append((pl, act))
            assert len(slots) <= 2, want
            if len(slots) > 0:
                p.out0_hi, p.out0_lo, p.out0_act = _ptr(slots[0][0][0]), _ptr(slots[0][0][1]), slots[0][1]
            if len(slots) > 1:
                p.out1_hi, p.out1_lo, p.out1_act = _ptr(slots[1][0][0]), _ptr(slots[1][0][1]), slots[1][1]
        p.passes = self.passes if a[1] is not None else 1
        # tensor cores: 64-channel K chunks; stride-1 convs also take cin % 8 == 0 (partly zero-filled last chunk)
        on_tc = pc.cin % 64 == 0 or (pc.cin % 8 == 0 and pc.stride == 1 and self.partial_chunk)
        p.impl = self.impl if on_tc else _lib.IMPL_SIMT
        if gn_groups > 0 and out.f32 is not None and pc.cout % gn_groups == 0:
            p.gn_groups = gn_groups
            rb, unit = ctypes.c_int32(0), ctypes.c_int32(0)
            if self.lib.mcq_conv_gn_layout(ctypes.byref(p), ctypes.byref(rb), ctypes.byref(unit)) == 0:
                part = torch.empty((x.n, rb.value, pc.cout // unit.value, 2), dtype=torch.float32, device=dev)
                p.gn_partials = _ptr(part)
                out.gn = (part, rb.value, unit.value)
                keep.append(part)
        info = {"flops": 2.0 * x.n * (x.h // pc.stride) * (x.w // pc.stride) * pc.cout * pc.cin * pc.ksize ** 2,
                "passes": p.passes, "impl": p.impl, "shape": (x.n, x.h, x.w, pc.cin, pc.cout, pc.ksize, pc.stride),
                # epilogue kind of the launch (profiling only): mode, fp32 output, residual / aux operands, plane pairs
                "epi": f"m{mode}{'F' if out.f32 is not None else ''}{'R' if res1 is not None else ''}"
                       f"{'r' if res2 is not None else ''}{'A' if aux is not None else ''}"
                       f"P{sum(t is not None for t in (out.raw, out.silu, out.sq))}"}
        # everything the launch touches stays referenced until it has been issued (a freed block could otherwise be
        # handed to a later layer of the same chain, whose clusters do not run in lock step)
        item = (p, (keep, out, pc, into), info)
        if (self.chain and p.impl == _lib.IMPL_TCGEN05 and out.gn is None and pc.cin % 64 == 0
                and (x.h // pc.stride) * (x.w // pc.stride) <= self.CHAIN_MAX_PIXELS):
            self._pending.append(item)
        else:
            self.flush()
            self._launch_one(item)
        return out

    # ------------------------------------------------------------------ blocks
    @staticmethod
    def needs_of(mod: nn.Module) -> Set[str]:
        """Representations of its input a module reads."""
        if isinstance(mod, ResidualBlock) and mod._skip is not None:
            return {"raw", "silu"}
        if isinstance(mod, (ResidualBlock, AttentionBlock)):
            return {"f32", "silu"}
        if isinstance(mod, (ResidualBlockWithStride, ResidualBlockShuffle)):
            return {"raw", "silu"}
        if isinstance(mod, (nn.Conv2d, nn.Sequential)):
            return {"raw"}
        raise NotImplementedError(f"mcquic_b200: no accelerated path for {type(mod).__name__}")

    def groupnorm(self, norm: nn.GroupNorm, x: Act, want: Set[str]) -> Act:
        """nn.GroupNorm on the fp32 output of a conv -> the representations the next conv reads (one launch)."""
        self.flush()
        if x.f32 is None or norm.num_channels != x.c:
            raise RuntimeError("mcquic_b200: GroupNorm expects the fp32 NHWC output of the producing convolution")
        if not norm.affine:
            raise NotImplementedError("mcquic_b200: GroupNorm without affine parameters is not on the accelerated path")
        dev = x.f32.device
        out = Act(x.n, x.h, x.w, x.c)
        if "f32" in want:
            out.f32 = torch.empty_like(x.f32)
        pl, act = (None, None), _lib.ACT_NONE
        planes = [name for name in ("raw", "silu", "sq") if name in want]
        if len(planes) > 1:
            raise NotImplementedError("mcquic_b200: GroupNorm writes one plane pair")
        if planes:
            pl = self._planes(x.n, x.h, x.w, x.c, dev)
            setattr(out, planes[0], pl)
            act = {"raw": _lib.ACT_NONE, "silu": _lib.ACT_SILU, "sq": _lib.ACT_SQUARE}[planes[0]]
        gamma = norm.weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        beta = norm.bias.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.gn is not None:
            # statistics came with the convolution: finalize (tiny) + ONE streaming normalise / split pass
            part, rb, unit = x.gn
            stats = torch.empty((x.n, norm.num_groups, 2), dtype=torch.float32, device=dev)
            with self._prof("mcq_groupnorm_apply"):
                _lib.check(self.lib.mcq_groupnorm_apply(_ptr(x.f32), _ptr(part), rb, unit, x.n, x.h, x.w, x.c,
                                                        norm.num_groups, _ptr(gamma), _ptr(beta), float(norm.eps),
                                                        _ptr(stats), _ptr(out.f32), _ptr(pl[0]), _ptr(pl[1]), act,
                                                        self._stream()), "mcq_groupnorm_apply")
            return out
        with self._prof("mcq_groupnorm"):
            _lib.check(self.lib.mcq_groupnorm(_ptr(x.f32), x.n, x.h, x.w, x.c, norm.num_groups, _ptr(gamma), _ptr(beta),
                                              float(norm.eps), _ptr(out.f32), _ptr(pl[0]), _ptr(pl[1]), act,
                                              self._stream()), "mcq_groupnorm")
        return out

    def residual_block(self, mod: ResidualBlock, x: Act, want: Set[str], res2: Optional[torch.Tensor] = None,
                       into: Optional[Act] = None) -> Act:
        identity = x.f32
        if mod._skip is not None:       # channel-changing block: conv1x1 on the un-activated x (blocks.py:73-75)
            identity = self.conv(self._packed_for(mod._skip), x.raw, x, {"f32"}).f32
        if isinstance(mod._branch[2], nn.GroupNorm):
            t = self.conv(self._packed_for(mod._branch[1]), x.silu, x, {"f32"}, gn_groups=mod._branch[2].num_groups)
            t = self.groupnorm(mod._branch[2], t, {"raw"})
            a = t.raw
        else:
            t = self.conv(self._packed_for(mod._branch[1]), x.silu, x, {"silu"})
            a = t.silu
        return self.conv(self._packed_for(mod._branch[3]), a, t, want, res1=identity, res2=res2, into=into)

    def residual_block_stride(self, mod: ResidualBlockWithStride, x: Act, want: Set[str]) -> Act:
        u = self.conv(self._packed_for(mod._branch[1]), x.silu, x, {"f32", "sq"})
        t = self.conv(self._packed_for(mod._branch[2]), u.sq, u, {"raw"}, mode=_lib.EPI_GDN, aux=u.f32,
                      a_scale=_lib.SQUARE_SCALE)
        s = self.conv(self._packed_for(mod._skip), x.raw, x, {"f32"})
        return self.conv(self._packed_for(mod._branch[3]), t.raw, t, want, res1=s.f32)

    def residual_block_shuffle(self, mod: ResidualBlockShuffle, x: Act, want: Set[str]) -> Act:
        sh = _lib.STORE_SHUFFLE_NHWC
        u = self.conv(self._packed_for(mod._branch[1][0], sh), x.silu, x, {"f32", "sq"})
        t = self.conv(self._packed_for(mod._branch[2]), u.sq, u, {"raw"}, mode=_lib.EPI_IGDN, aux=u.f32,
                      a_scale=_lib.SQUARE_SCALE)
        s = self.conv(self._packed_for(mod._skip[0], sh), x.raw, x, {"f32"})
        return self.conv(self._packed_for(mod._branch[3]), t.raw, t, want, res1=s.f32)

    def attention_block(self, mod: AttentionBlock, x: Act, want: Set[str]) -> Act:
        def main_branch():
            a = x
            for i in range(3):
                a = self.residual_block(mod._mainBranch[i], a, {"f32", "silu"} if i < 2 else {"f32"})
            return a

        def side_branch():
            b = x
            for i in range(3):
                b = self.residual_block(mod._sideBranch[i], b, {"f32", "silu"} if i < 2 else {"raw"})
            return b

        dense = isinstance(mod._mainBranch[0]._branch[2], nn.GroupNorm)   # its GroupNorm launches cannot join a chain
        a, b = self.parallel(main_branch, side_branch, chain=self.can_chain(x) and not dense)
        return self.conv(self._packed_for(mod._sideBranch[3]), b.raw, b, want, mode=_lib.EPI_GATE, res1=x.f32,
                         aux=a.f32)

    def run(self, mod: nn.Module, x: Act, want: Set[str], tail: Optional[Tuple[torch.Tensor, float]] = None,
            into: Optional[Act] = None) -> Act:
        """tail = (tensor, scale): an extra fp32 NHWC term added by the module's last epilogue.
        into (ResidualBlock and the final pixel-shuffle conv only): caller-owned output buffers, see conv()."""
        if isinstance(mod, ResidualBlock):
            assert tail is None or tail[1] == 1.0
            return self.residual_block(mod, x, want, None if tail is None else tail[0], into=into)
        if into is not None and not (isinstance(mod, nn.Sequential) and len(mod) == 2):
            raise NotImplementedError("mcquic_b200: `into` is supported for ResidualBlock and the final conv only")
        if isinstance(mod, nn.Conv2d):
            if tail is None:
                return self.conv(self._packed_for(mod), x.raw, x, want)
            return self.conv(self._packed_for(mod), x.raw, x, want, res1=tail[0], res1_scale=tail[1])
        assert tail is None
        if isinstance(mod, ResidualBlockWithStride):
            return self.residual_block_stride(mod, x, want)
        if isinstance(mod, ResidualBlockShuffle):
            return self.residual_block_shuffle(mod, x, want)
        if isinstance(mod, AttentionBlock):
            return self.attention_block(mod, x, want)
        if isinstance(mod, nn.Sequential) and len(mod) == 2 and isinstance(mod[1], nn.PixelShuffle):
            # pixelShuffle3x3 used stand-alone = last decoder layer -> fp32 NCHW pixels
            return self.conv(self._packed_for(mod[0], _lib.STORE_SHUFFLE_NCHW), x.raw, x, want, into=into)
        raise NotImplementedError(f"mcquic_b200: no accelerated path for {type(mod).__name__}")

    def run_seq(self, mods: Sequence[nn.Module], x: Act, want: Set[str],
                tail: Optional[Tuple[torch.Tensor, float]] = None) -> Act:
        mods = list(mods)
        for i, mod in enumerate(mods):
            last = i + 1 == len(mods)
            x = self.run(mod, x, want if last else self.needs_of(mods[i + 1]), tail if last else None)
        return x

    # ------------------------------------------------------------------ boundary ops
    def _packed_stem(self, conv: nn.Conv2d):
        """tcgen05 operand of the stem: fp16 [cout_pad, 64], row = [w_lo (27 + 5 zeros) | w_hi (27 + 5 zeros)] of
        w * 2^e (csrc/stem_tc.cuh), plus (2^-e, bias, cout_pad); cached per weight version"""
        key = (id(conv), "stem")
        try:
            ver = (conv.weight._version, conv.bias._version, conv.weight.data_ptr(), conv.bias.data_ptr())
        except RuntimeError:
            ver = (0, 0, conv.weight.data_ptr(), conv.bias.data_ptr())
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver and hit[2]() is conv:
            return hit[1]
        cout = conv.out_channels
        hi, lo, scale = split_weight(conv.weight.detach().float().reshape(cout, 27))
        cout_pad = (cout + 15) // 16 * 16
        rows = torch.zeros((cout_pad, 64), dtype=torch.float16, device=conv.weight.device)
        rows[:cout, 0:27] = lo
        rows[:cout, 32:59] = hi
        packed = (rows.contiguous(), scale, conv.bias.detach().contiguous().float(), cout_pad)
        self._packed[key] = (ver, packed, weakref.ref(conv))
        return packed

    def stem(self, conv: nn.Conv2d, x: torch.Tensor, pad: Tuple[int, int, int, int], want: Set[str],
             into: Optional[Act] = None) -> Act:
        """conv3x3 s2 on the NCHW image; pad = (top, left, padded_h, padded_w) (AlignedPadding folded in).
        x: fp32 in [-1, 1], or uint8 (the reference's input transform convert_image_dtype + (x - 0.5) * 2, demo.py:110-118,
        is then applied inside the kernel).  tcgen05 kernel (fp32-grade 3-pass) for cout <= 128, FFMA kernel otherwise /
        for impl='simt'.  (Chunked execution calls this on batch slices of the image tensor.)"""
        self.flush()
        n, c, h, w = x.shape
        if c != 3 or conv.in_channels != 3 or conv.stride[0] != 2:
            raise RuntimeError("mcquic_b200: the stem expects [n, 3, h, w] input and conv3x3(3, C, stride=2)")
        if x.dtype not in (torch.float32, torch.uint8):
            raise RuntimeError(f"mcquic_b200: images must be float32 in [-1, 1] or uint8, got {x.dtype}")
        top, left, hp, wp = pad
        cout = conv.out_channels
        dev = x.device
        out = into if into is not None else Act(n, hp // 2, wp // 2, cout)
        if into is None:
            if "f32" in want:
                out.f32 = torch.empty((n, out.h, out.w, cout), dtype=torch.float32, device=dev)
            if "silu" in want:
                out.silu = self._planes(n, out.h, out.w, cout, dev)
            elif "raw" in want:
                out.raw = self._planes(n, out.h, out.w, cout, dev)
        pl, act = (None, None), _lib.ACT_NONE
        if "silu" in want:
            pl, act = out.silu, _lib.ACT_SILU
        elif "raw" in want:
            pl = out.raw
        xc = x.contiguous()
        u8 = 1 if x.dtype == torch.uint8 else 0
        f32 = out.f32 if "f32" in want else None
        if self.impl == _lib.IMPL_TCGEN05 and not self.emulated and cout <= 128 and cout % 8 == 0 and self.passes == 3 \
                and self.stem_tc:
            rows, scale, bias, cout_pad = self._packed_stem(conv)
            with self._prof("mcq_stem_conv_tc"):
                _lib.check(self.lib.mcq_stem_conv_tc(_ptr(xc), u8, n, h, w, top, left, hp, wp, _ptr(rows), scale,
                                                     _ptr(bias), cout, cout_pad, _ptr(f32), _ptr(pl[0]), _ptr(pl[1]),
                                                     act, self._stream()), "mcq_stem_conv_tc")
            return out
        wgt = conv.weight.detach().reshape(cout, 27).contiguous().float()
        bias = conv.bias.detach().contiguous().float()
        with self._prof("mcq_stem_conv"):
            _lib.check(self.lib.mcq_stem_conv(_ptr(xc), u8, n, h, w, top, left, hp, wp, _ptr(wgt), _ptr(bias), cout,
                                              _ptr(f32), _ptr(pl[0]), _ptr(pl[1]), act, self._stream()),
                       "mcq_stem_conv")
        return out

    def add_scaled(self, x: torch.Tensor, y: torch.Tensor, alpha: float, shape: Tuple[int, int, int, int],
                   want: Set[str]) -> Act:
        """x + alpha * y on fp32 NHWC tensors of `shape` = (n, h, w, c) -> the representations in `want`."""
        self.flush()
        n, h, w, c = shape
        if tuple(x.shape) != tuple(shape) or tuple(y.shape) != tuple(shape):
            raise RuntimeError(f"mcquic_b200: add_scaled operands must both be {tuple(shape)}")
        out = Act(n, h, w, c)
        if "f32" in want:
            out.f32 = torch.empty_like(x)
        planes = [name for name in ("raw", "silu", "sq") if name in want]
        if len(planes) > 1:
            raise NotImplementedError("mcquic_b200: add_scaled writes one plane pair")
        pl, act = (None, None), _lib.ACT_NONE
        if planes:
            pl = self._planes(n, h, w, c, x.device)
            setattr(out, planes[0], pl)
            act = {"raw": _lib.ACT_NONE, "silu": _lib.ACT_SILU, "sq": _lib.ACT_SQUARE}[planes[0]]
        with self._prof("mcq_add_scaled"):
            _lib.check(self.lib.mcq_add_scaled(_ptr(x), _ptr(y), float(alpha), x.numel(), _ptr(out.f32), _ptr(pl[0]),
                                               _ptr(pl[1]), act, self._stream()), "mcq_add_scaled")
        return out

    def from_nchw(self, x: torch.Tensor, want: Set[str], pad_channels_to: int = 1) -> Act:
        """pad_channels_to: zero channels are appended up to a multiple of it (RGB -> 8 for Neon's first conv)."""
        self.flush()
        if x.shape[1] % pad_channels_to != 0:
            extra = pad_channels_to - x.shape[1] % pad_channels_to
            x = torch.cat([x, x.new_zeros(x.shape[0], extra, x.shape[2], x.shape[3])], 1)
        n, c, h, w = x.shape
        out = Act(n, h, w, c)
        xc = x.contiguous().float()
        dev = x.device
        if "f32" in want:
            out.f32 = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
        slots = []
        for name, act in (("raw", _lib.ACT_NONE), ("silu", _lib.ACT_SILU), ("sq", _lib.ACT_SQUARE)):
            if name in want:
                pl = self._planes(n, h, w, c, dev)
                setattr(out, name, pl)
                slots.append((pl, act))
        slots += [((None, None), 0)] * (2 - len(slots))
        with self._prof("mcq_nchw_to_nhwc"):
            _lib.check(self.lib.mcq_nchw_to_nhwc(_ptr(xc), n, c, h, w, _ptr(out.f32), _ptr(slots[0][0][0]),
                                                 _ptr(slots[0][0][1]), slots[0][1], _ptr(slots[1][0][0]),
                                                 _ptr(slots[1][0][1]), slots[1][1], self._stream()), "mcq_nchw_to_nhwc")
        return out

    def to_nchw(self, x: Act) -> torch.Tensor:
        self.flush()
        out = torch.empty((x.n, x.c, x.h, x.w), dtype=torch.float32, device=x.f32.device)
        with self._prof("mcq_nhwc_to_nchw"):
            _lib.check(self.lib.mcq_nhwc_to_nchw(_ptr(x.f32), x.n, x.c, x.h, x.w, _ptr(out), self._stream()),
                       "mcq_nhwc_to_nchw")
        return out

    def run_seq_nchw(self, mods: Sequence[nn.Module], x: torch.Tensor) -> torch.Tensor:
        """A chain of blocks / convs on an NCHW fp32 tensor (the values-only `forward` paths of the quantizers)."""
        if not x.is_cuda and not self.emulated:
            raise RuntimeError("mcquic_b200 runs on CUDA tensors only (no CPU fallback)")
        mods = list(mods)
        with torch.no_grad():
            out = self.run_seq(mods, self.from_nchw(x, self.needs_of(mods[0])), {"f32"})
            self.flush()
            return self.to_nchw(out)

    def run_module_nchw(self, mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
        """Stand-alone execution of one block on an NCHW tensor (API parity with calling the reference module)."""
        if not x.is_cuda and not self.emulated:
            raise RuntimeError("mcquic_b200 runs on CUDA tensors only (no CPU fallback)")
        with torch.no_grad():
            act = self.from_nchw(x, self.needs_of(mod))
            return self.to_nchw(self.run(mod, act, {"f32"}))

    # ------------------------------------------------------------------ VQ
    def vq_assign(self, x_nhwc: torch.Tensor, codebook: torch.Tensor, c2: torch.Tensor, n: int, h: int, w: int,
                  logits: bool = False, logit_scale: Optional[torch.Tensor] = None,
                  hist: Optional[torch.Tensor] = None, packed=None):
        """packed = pack_codebook(codebook): split-fp16 codebook for the tensor-core paths.
        Routing: fused tcgen05 kernel (d in {32, 64}: hard codes and soft logits in one launch) -> GEMM-epilogue
        argmin on the conv kernel (d % 64 == 0, hard codes only, needs `packed`) -> SIMT kernel (any shape)."""
        self.flush()
        m, k, d = codebook.shape
        dev = x_nhwc.device
        codes = torch.empty((n, m, h, w), dtype=torch.int64, device=dev)
        on_tc = not self.emulated and self.impl == _lib.IMPL_TCGEN05
        if on_tc and self.lib.mcq_vq_fused_supported(h, w, k, d):
            if packed is None or len(packed) < 4:
                packed = pack_codebook(codebook)
            lg = torch.empty((n, m, h, w, k), dtype=torch.float32, device=dev) if logits else None
            with self._prof("mcq_vq_assign_fused"):
                _lib.check(self.lib.mcq_vq_assign_fused(_ptr(x_nhwc), _ptr(packed[3]), packed[2], _ptr(c2), _ptr(codes),
                                                        _ptr(lg), _ptr(logit_scale), _ptr(hist), n, h, w, m, k, d,
                                                        self._stream()), "mcq_vq_assign_fused")
            return (codes, lg) if logits else codes
        use_tc = packed is not None and not logits and d % 64 == 0 and k % 32 == 0 and on_tc
        if use_tc:
            nbytes = int(self.lib.mcq_vq_workspace_bytes(n, h, w, m, k, d))
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            with self._prof("mcq_vq_assign_tc"):
                _lib.check(self.lib.mcq_vq_assign_tc(_ptr(x_nhwc), _ptr(packed[0]), _ptr(packed[1]), packed[2], _ptr(c2),
                                                     _ptr(codes), _ptr(hist), n, h, w, m, k, d,
                                                     ctypes.c_void_p(ws.data_ptr() + off), nbytes, self._stream()),
                           "mcq_vq_assign_tc")
            return codes
        lg = torch.empty((n, m, h, w, k), dtype=torch.float32, device=dev) if logits else None
        with self._prof("mcq_vq_assign"):
            _lib.check(self.lib.mcq_vq_assign(_ptr(x_nhwc), _ptr(codebook), _ptr(c2), _ptr(codes), _ptr(lg),
                                              _ptr(logit_scale), _ptr(hist), n, h, w, m, k, d, self._stream()),
                       "mcq_vq_assign")
        return (codes, lg) if logits else codes

    def vq_dequant(self, codes: torch.Tensor, codebook: torch.Tensor, want: Set[str],
                   status: Optional[torch.Tensor] = None) -> Act:
        self.flush()
        n, m, h, w = codes.shape
        _, k, d = codebook.shape
        out = Act(n, h, w, m * d)
        dev = codes.device
        if "f32" in want:
            out.f32 = torch.empty((n, h, w, m * d), dtype=torch.float32, device=dev)
        slots = []
        for name, act in (("raw", _lib.ACT_NONE), ("silu", _lib.ACT_SILU)):
            if name in want:
                pl = self._planes(n, h, w, m * d, dev)
                setattr(out, name, pl)
                slots.append((pl, act))
        slots += [((None, None), 0)] * (2 - len(slots))
        with self._prof("mcq_vq_dequant"):
            _lib.check(self.lib.mcq_vq_dequant(_ptr(codes), _ptr(codebook), n, h, w, m, k, d, _ptr(out.f32),
                                               _ptr(slots[0][0][0]), _ptr(slots[0][0][1]), slots[0][1],
                                               _ptr(slots[1][0][0]), _ptr(slots[1][0][1]), slots[1][1], _ptr(status),
                                               self._stream()), "mcq_vq_dequant")
        return out

    def code_histogram(self, codes: List[torch.Tensor], ks: Sequence[int], out: Optional[torch.Tensor] = None):
        """One flat int32 buffer [sum_l m*k_l] (level-major) -- the only thing that ever crosses NVLink."""
        self.flush()
        m = codes[0].shape[1]
        total = sum(m * k for k in ks)
        if out is None:
            out = torch.zeros(total, dtype=torch.int32, device=codes[0].device)
        off = 0
        for code, k in zip(codes, ks):
            n, mm, h, w = code.shape
            view = out[off:off + mm * k]
            with self._prof("mcq_code_histogram"):
                _lib.check(self.lib.mcq_code_histogram(_ptr(code.contiguous()), n, mm, h * w, k, _ptr(view),
                                                       self._stream()), "mcq_code_histogram")
            off += mm * k
        return out


_DEFAULT: Optional[Engine] = None


def default_engine() -> Engine:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Engine()
    return _DEFAULT
