"""Host-side entropy coding of code maps (SURVEY.md section 8f NEXT-1): ctypes binding of libmcquic_entropy.so
(include/mcquic_entropy.h) -- a multi-threaded C++ rANS coder whose byte streams are bit-identical to the reference's
`mcquic.rans` extension -- plus the reference's `.mcq` header dataclasses (mcquic/utils/specification.py:56-160).
The entropy coder stays on the host (north_star); only code maps cross PCIe."""
import ctypes
import os
import shutil
import subprocess
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np
import torch

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "entropy")
_LIB_PATH = os.path.join(_HERE, "libmcquic_entropy.so")
_SRC = os.path.join(_HERE, "mcq_entropy.cpp")
_LIB = None

_i32, _i64, _p = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
SYMBOLS = {
    "mcq_pmf_to_quantized_cdf": (ctypes.c_int, [_p, _i32, _p]),
    "mcq_rans_stream_capacity": (_i64, [_i64]),
    "mcq_rans_encode_level": (ctypes.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, _i64, _p, _i32]),
    "mcq_rans_decode_level": (ctypes.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _i32]),
    "mcq_entropy_version": (ctypes.c_int, []),
}


def build(force: bool = False) -> str:
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_SRC) > os.path.getmtime(_LIB_PATH)
    if force or stale:
        gxx = shutil.which("g++")
        if gxx is None:
            raise RuntimeError("g++ not found: cannot build libmcquic_entropy.so")
        tmp = f"{_LIB_PATH}.{os.getpid()}.tmp"      # concurrent ranks must never dlopen a half-written library
        res = subprocess.run([gxx, "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", tmp, _SRC],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stderr)
        os.replace(tmp, _LIB_PATH)
    return _LIB_PATH


def load():
    global _LIB
    if _LIB is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB_PATH):
                raise
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def pmf_to_quantized_cdf(pmf: Sequence[float]) -> np.ndarray:
    """pmfToQuantizedCDF(pmf, 16) of the reference (ops.cpp:42-111): uint32 [k + 1]."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.empty(p.shape[0] + 1, dtype=np.uint32)
    if load().mcq_pmf_to_quantized_cdf(p.ctypes.data, p.shape[0], cdf.ctypes.data) != 0:
        raise ValueError("Invalid `pmf`: negative / non-finite element or all-zero")   # std::domain_error upstream
    return cdf


def encode_level(codes: torch.Tensor, cdfs: np.ndarray, threads: int = 0) -> List[bytes]:
    """codes int64 [n, m, h, w] (any device) + cdfs uint32 [m, k+1] -> one rANS stream per image."""
    c = codes.detach().to("cpu", torch.int64).contiguous()
    n, m, h, w = c.shape
    k = cdfs.shape[1] - 1
    cap = int(load().mcq_rans_stream_capacity(m * h * w))
    out = np.empty((n, cap), dtype=np.uint8)
    sizes = np.empty(n, dtype=np.int32)
    cd = np.ascontiguousarray(cdfs, dtype=np.uint32)
    rc = load().mcq_rans_encode_level(c.data_ptr(), n, m, h * w, k, cd.ctypes.data, out.ctypes.data, cap,
                                      sizes.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"rANS encode failed (code {rc}): code index outside its codebook or zero-probability symbol")
    return [out[i, :sizes[i]].tobytes() for i in range(n)]


MAX_SYMBOLS_PER_STREAM = 1 << 26      # 64 Mi codes per image and level (an 8192 x 8192 code map): header sanity bound


def decode_level(streams: Sequence[bytes], m: int, h: int, w: int, cdfs: np.ndarray, threads: int = 0) -> torch.Tensor:
    """inverse of encode_level.  m, h, w usually come from an untrusted `.mcq` header: they are checked against the CDF
    table and a sanity bound before anything is allocated, and the C++ decoder stops with an error when a stream runs
    out of words (truncated file / header claiming more symbols than the stream holds)."""
    n = len(streams)
    if n < 1:
        raise RuntimeError("rANS decode: no streams")
    if cdfs.ndim != 2 or m != cdfs.shape[0]:
        raise RuntimeError(f"rANS decode: header says m = {m}, the model's CDF table has {cdfs.shape[0]} codebooks")
    if h < 1 or w < 1 or m * h * w > MAX_SYMBOLS_PER_STREAM:
        raise RuntimeError(f"rANS decode: implausible code map {m} x {h} x {w}")
    k = cdfs.shape[1] - 1
    stride = max(len(s) for s in streams)
    stride = (stride + 3) // 4 * 4
    buf = np.zeros((n, stride), dtype=np.uint8)
    sizes = np.empty(n, dtype=np.int32)
    for i, s in enumerate(streams):
        buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        sizes[i] = len(s)
    out = torch.empty((n, m, h, w), dtype=torch.int64)
    cd = np.ascontiguousarray(cdfs, dtype=np.uint32)
    rc = load().mcq_rans_decode_level(buf.ctypes.data, sizes.ctypes.data, stride, n, m, h * w, k, cd.ctypes.data,
                                      out.data_ptr(), threads)
    if rc != 0:
        raise RuntimeError(f"rANS decode failed (code {rc}): malformed stream or CDF")
    return out


# ---- header records of the reference's container (mcquic/utils/specification.py:56-119)
@dataclass
class ImageSize:
    height: int
    width: int
    channel: int

    @property
    def Pixels(self) -> int:
        return self.height * self.width


@dataclass
class CodeSize:
    m: List[int]
    heights: List[int]
    widths: List[int]
    k: List[int]


@dataclass
class FileHeader:
    version: str
    qp: str
    codeSize: CodeSize
    imageSize: ImageSize

    @property
    def CodeSize(self) -> CodeSize:
        return self.codeSize

    @property
    def ImageSize(self) -> ImageSize:
        return self.imageSize


def bpp(binaries: Sequence[bytes], image: ImageSize) -> float:
    """File.BPP of the reference (specification.py:158-160)."""
    return sum(len(b) for b in binaries) * 8 / image.Pixels
