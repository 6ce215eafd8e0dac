"""Pure-Python restatement of the reference's rANS entropy coder -- TEST INFRASTRUCTURE ONLY (small cases; Python ints).

Follows third_party/CompressAI/cpp_exts/ops.cpp:42-111 (pmfToQuantizedCDF), buffered_rans_encoder.cpp:104-196
(symbol buffering, reverse encoding, flush), rans_decoder.cpp:104-173 and ryg_rans/rans64.h:59-140 (64-bit state,
32-bit renormalisation, L = 2^31), without the bypass path (never taken for codes in [0, k)).  Pinned bit-for-bit
against the reference's compiled module (oracle/_ref, oracle/build_ref.py) in tests/test_entropy.py."""
import struct
from typing import List, Sequence

import numpy as np

PRECISION = 16
RANS_L = 1 << 31


def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = PRECISION) -> List[int]:
    pmf32 = np.asarray(pmf, dtype=np.float32)
    if (pmf32 < 0).any() or not np.isfinite(pmf32).all():
        raise ValueError("Invalid `pmf`")
    scaled = pmf32 * np.float32(1 << precision)
    # std::round: half away from zero
    cdf = [0] + [int(np.floor(float(v) + 0.5)) for v in scaled]
    total = sum(cdf) & 0xFFFFFFFF
    if total == 0:
        raise ValueError("Invalid `pmf`")
    cdf = [((1 << precision) * p) // total for p in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best, steal = None, -1
            for j in range(len(cdf) - 1):
                f = cdf[j + 1] - cdf[j]
                if f > 1 and (best is None or f < best):
                    best, steal = f, j
            assert steal != -1
            if steal < i:
                for j in range(steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, steal + 1):
                    cdf[j] += 1
    return cdf


def encode(symbols: Sequence[int], indexes: Sequence[int], cdfs: Sequence[Sequence[int]]) -> bytes:
    x = RANS_L
    words: List[int] = []          # emitted last-to-first
    for s, ci in zip(reversed(list(symbols)), reversed(list(indexes))):
        cdf = cdfs[ci]
        start, freq = cdf[s], cdf[s + 1] - cdf[s]
        x_max = ((RANS_L >> PRECISION) << 32) * freq
        if x >= x_max:
            words.append(x & 0xFFFFFFFF)
            x >>= 32
        x = ((x // freq) << PRECISION) + (x % freq) + start
    words.append(x >> 32)
    words.append(x & 0xFFFFFFFF)
    return b"".join(struct.pack("<I", w) for w in reversed(words))


def decode(stream: bytes, indexes: Sequence[int], cdfs: Sequence[Sequence[int]]) -> List[int]:
    words = struct.unpack(f"<{len(stream) // 4}I", stream)
    x = words[0] | (words[1] << 32)
    pos = 2
    out = []
    mask = (1 << PRECISION) - 1
    for ci in indexes:
        cdf = cdfs[ci]
        cum = x & mask
        s = next(i for i, v in enumerate(cdf) if v > cum) - 1
        x = (cdf[s + 1] - cdf[s]) * (x >> PRECISION) + (x & mask) - cdf[s]
        if x < RANS_L:
            x = (x << 32) | words[pos]
            pos += 1
        out.append(s)
    return out
