"""NEXT-1 row (host entropy coder): mcquic_b200's C++ rANS library vs (1) the reference's OWN compiled coder
(oracle/_ref, built from /root/reference by oracle/build_ref.py; travels to the GPU box as a prebuilt file),
(2) the pure-Python restatement oracle/rans_oracle.py, (3) known answers: byte counts / bpp of the reference flow at
qp=1 under the uniform prior (424 + 96 + 24 bytes = 0.0664 bpp at 256x256)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from common import golden_codes, load_golden
from mcquic_b200 import entropy
from oracle import build_ref, rans_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = build_ref.load()
needs_ref = pytest.mark.skipif(REF is None, reason="reference coder (oracle/_ref) not built and /root/reference absent")


def _ref_encode(code_img, cdfs, k):
    m, h, w = code_img.shape
    idx = torch.arange(m)[:, None, None].expand(m, h, w).flatten().int().tolist()     # entropyCoder.py:114-118
    return REF.RansEncoder().encodeWithIndexes(code_img.flatten().int().tolist(), idx, [c.tolist() for c in cdfs],
                                               [k + 2] * m, [0] * (m * h * w))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mcquic_entropy.h")).read()
    declared = set(re.findall(r"\b(mcq_[a-z0-9_]+)\s*\(", header))
    assert declared == set(entropy.SYMBOLS)
    lib = entropy.load()
    assert all(hasattr(lib, s) for s in declared) and lib.mcq_entropy_version() == 1


@needs_ref
@pytest.mark.parametrize("k", [2, 16, 512, 2048, 8192])
def test_quantized_cdf_matches_reference(k):
    rng = np.random.default_rng(k)
    for trial in range(4):
        pmf = rng.random(k).astype(np.float32) ** (1 + 3 * trial)            # increasingly peaked
        if trial >= 2 and k > 4:
            pmf[rng.integers(0, k, k // 3)] = 0                                # never-used codewords: frequency stealing
        pmf /= pmf.sum()
        mine = entropy.pmf_to_quantized_cdf(pmf)
        assert mine.tolist() == list(REF.pmfToQuantizedCDF(pmf.tolist(), 16))
        assert mine[0] == 0 and mine[-1] == 65536 and (np.diff(mine.astype(np.int64)) > 0).all()
    uniform = np.full(k, 1.0 / k, dtype=np.float32)                            # the initial prior (entropyCoder.py:22)
    assert entropy.pmf_to_quantized_cdf(uniform).tolist() == list(REF.pmfToQuantizedCDF(uniform.tolist(), 16))


def test_quantized_cdf_against_python_oracle_and_errors():
    rng = np.random.default_rng(3)
    for k in (3, 40, 300):
        pmf = rng.random(k).astype(np.float32)
        pmf[::7] = 0
        pmf /= pmf.sum()
        assert entropy.pmf_to_quantized_cdf(pmf).tolist() == rans_oracle.pmf_to_quantized_cdf(pmf)
    for bad in ([0.5, -0.1, 0.6], [0.0, 0.0], [float("nan"), 1.0]):
        with pytest.raises(ValueError):
            entropy.pmf_to_quantized_cdf(bad)


@needs_ref
@pytest.mark.parametrize("n,m,h,w,k", [(3, 1, 16, 16, 8192), (2, 2, 8, 8, 512), (4, 6, 4, 4, 2048), (2, 3, 5, 7, 100)])
def test_streams_bit_identical_to_reference_coder(n, m, h, w, k):
    rng = np.random.default_rng(n * 1000 + k)
    cdfs = np.stack([entropy.pmf_to_quantized_cdf(rng.random(k) ** 2) for _ in range(m)])
    codes = torch.from_numpy(rng.integers(0, k, (n, m, h, w)))
    mine = entropy.encode_level(codes, cdfs)
    for i in range(n):
        assert mine[i] == _ref_encode(codes[i], cdfs, k)
    assert torch.equal(entropy.decode_level(mine, m, h, w, cdfs), codes)       # our decoder
    idx = torch.arange(m)[:, None, None].expand(m, h, w).flatten().int().tolist()
    back = REF.RansDecoder().decodeWithIndexes(mine[0], idx, [c.tolist() for c in cdfs], [k + 2] * m, [0] * (m * h * w))
    assert back == codes[0].flatten().tolist()                                  # the reference's decoder reads ours


def test_streams_match_python_oracle_and_round_trip():
    rng = np.random.default_rng(11)
    m, h, w, k = 2, 3, 4, 37
    cdfs = np.stack([entropy.pmf_to_quantized_cdf(rng.random(k)) for _ in range(m)])
    codes = torch.from_numpy(rng.integers(0, k, (3, m, h, w)))
    mine = entropy.encode_level(codes, cdfs, threads=2)
    idx = [j // (h * w) for j in range(m * h * w)]
    for i in range(3):
        flat = codes[i].flatten().tolist()
        assert mine[i] == rans_oracle.encode(flat, idx, [c.tolist() for c in cdfs])
        assert rans_oracle.decode(mine[i], idx, [c.tolist() for c in cdfs]) == flat
    assert torch.equal(entropy.decode_level(mine, m, h, w, cdfs), codes)
    # one-symbol stream (the reference coder itself overruns its output buffer below ~3 symbols, so no _ref here)
    one = torch.tensor([[[[1]]]])
    tiny = entropy.encode_level(one, cdfs[:1])
    assert tiny[0] == rans_oracle.encode([1], [0], [cdfs[0].tolist()]) and len(tiny[0]) == 8
    assert torch.equal(entropy.decode_level(tiny, 1, 1, 1, cdfs[:1]), one)
    with pytest.raises(RuntimeError):
        entropy.encode_level(torch.full((1, m, h, w), k), cdfs)                # code outside [0, k)
    with pytest.raises(RuntimeError):
        entropy.decode_level([b"\x00" * 6], m, h, w, cdfs)                     # truncated stream


def test_decoder_never_reads_past_a_stream_and_headers_are_validated():
    """the symbol count comes from an untrusted `.mcq` header.  A valid 12-byte stream decoded with
    h = w = 4096 used to walk off the buffer (segfault); now the C++ decoder stops with -3 when the words run out, and the
    Python layer checks header fields against the model before the coder sees them."""
    rng = np.random.default_rng(5)
    k = 64
    cdfs = np.stack([entropy.pmf_to_quantized_cdf(rng.random(k))])
    codes = torch.from_numpy(rng.integers(0, k, (1, 1, 2, 2)))
    stream = entropy.encode_level(codes, cdfs)[0]
    assert len(stream) <= 16
    assert torch.equal(entropy.decode_level([stream], 1, 2, 2, cdfs), codes)
    for h, w in ((4096, 4096), (64, 64), (3, 3)):                      # header claims more symbols than the stream holds
        with pytest.raises(RuntimeError, match="rANS decode"):
            entropy.decode_level([stream], 1, h, w, cdfs)
    big = entropy.encode_level(torch.from_numpy(rng.integers(0, k, (1, 1, 32, 32))), cdfs)[0]
    for cut in (8, 12, len(big) // 2 // 4 * 4, len(big) - 4):           # truncated files
        with pytest.raises(RuntimeError, match="rANS decode"):
            entropy.decode_level([big[:cut]], 1, 32, 32, cdfs)
    with pytest.raises(RuntimeError, match="rANS decode"):
        entropy.decode_level([stream], 2, 2, 2, cdfs)                  # m larger than the CDF table
    with pytest.raises(RuntimeError, match="implausible"):
        entropy.decode_level([stream], 1, 1 << 20, 1 << 20, cdfs)      # would be a multi-terabyte allocation
    with pytest.raises(RuntimeError):
        entropy.decode_level([], 1, 2, 2, cdfs)
    bad = cdfs.copy()
    bad[0, -1] = 60000                                                # table does not cover [0, 2^16)
    with pytest.raises(RuntimeError, match="rANS decode"):
        entropy.decode_level([stream], 1, 2, 2, bad)
    # CodeFrequency.decompress: header vs model
    from mcquic_b200 import Compressor
    model = Compressor(32, 2, [16, 8]).eval()
    coder = model._quantizer._entropyCoder
    cds = [torch.from_numpy(rng.integers(0, 16, (2, 2, 4, 4))), torch.from_numpy(rng.integers(0, 8, (2, 2, 2, 2)))]
    bins, sizes = coder.compress(cds)
    assert all(torch.equal(a, b) for a, b in zip(coder.decompress(bins, sizes), cds))
    def header(**kw):
        f = dict(m=[2, 2], heights=[4, 2], widths=[4, 2], k=[16, 8])
        f.update(kw)
        return [entropy.CodeSize(**f)] * 2
    for hd in (header(m=[3, 2]), header(k=[16, 16]), header(m=[2], heights=[4], widths=[4], k=[16]),
               header(heights=[4096, 2], widths=[4096, 2]), header(heights=[4, 2, 1]),
               [entropy.CodeSize([2, 2], [4, 2], [4, 2], [16, 8]), entropy.CodeSize([2, 2], [8, 2], [4, 2], [16, 8])]):
        with pytest.raises(RuntimeError):
            coder.decompress(bins, hd)
    with pytest.raises(RuntimeError):
        coder.decompress([b[:1] for b in bins], sizes)                 # a level's stream is missing
    with pytest.raises(RuntimeError):
        coder.decompress(bins[:1], sizes)


def test_compress_decompress_api_and_bpp_known_answer():
    """qp=1, 256x256, uniform prior: the reference flow gives 424 + 96 + 24 bytes = 0.0664 bpp.
    Uses the golden codes (reference outputs); compress()/decompress() run through the emulated engine on CPU."""
    from emulator import EmulatedLib
    from mcquic_b200 import Compressor
    from mcquic_b200.engine import Engine
    g, cfg = load_golden("compressor_qp1_256")
    ref_codes = golden_codes(g, 3)
    model = Compressor(cfg["channel"], cfg["m"], cfg["k"]).eval()
    binaries, sizes = model._quantizer._entropyCoder.compress(ref_codes)
    assert [len(b) for b in binaries[0]] == [424, 96, 24]
    assert entropy.bpp(binaries[0], entropy.ImageSize(256, 256, 3)) == pytest.approx(544 * 8 / 65536)
    assert sizes[0].m == [1, 1, 1] and sizes[0].heights == [16, 8, 4] and sizes[0].k == [8192, 2048, 512]
    back = model._quantizer._entropyCoder.decompress(binaries, sizes)
    assert all(torch.equal(a, b) for a, b in zip(back, ref_codes))
    if REF is not None:
        for lv, k in enumerate(cfg["k"]):
            cdf = np.array(REF.pmfToQuantizedCDF([1.0 / k] * k, 16), dtype=np.uint32)[None]
            assert binaries[0][lv] == _ref_encode(ref_codes[lv][0], cdf, k)
    # whole public API on an unaligned image: compress -> decompress crops back to the input size
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    small = Compressor(32, 2, [16, 8]).eval()
    small.load_state_dict(synthetic_state_dict(32, 2, [16, 8], seed=0))
    small._engine = Engine(lib=EmulatedLib())
    x = uniform((2, 3, 100, 130), "compress.image", 0)
    codes, bins, headers = small.compress(x)
    assert len(bins) == 2 and len(bins[0]) == 2 and headers[0].ImageSize.height == 100 and headers[0].ImageSize.width == 130
    out = small.decompress(bins, headers)
    assert tuple(out.shape) == (2, 3, 100, 130)
    full = small.decode(codes)                                                  # [2, 3, 128, 256]
    top, left = (128 - 100) // 2, (256 - 130) // 2
    assert torch.equal(out, full[..., top:top + 100, left:left + 130])
    # frequency update from a histogram changes the CDFs and the streams shrink for the seen codes
    hist = torch.cat([h.flatten() for h in __import__("oracle.mcquic_oracle", fromlist=["x"]).code_histogram(codes, [16, 8])]).int()
    before = sum(len(b) for img in bins for b in img)
    for _ in range(20):
        small._quantizer._entropyCoder.update(hist)
    after = sum(len(b) for img in small.compress(x)[1] for b in img)
    assert after <= before


def test_decode_tables_are_cached_by_cdf_contents_and_threads_are_optional():
    """the decoder keeps its cum_freq -> symbol tables across calls, keyed by the CDF's contents: more distinct CDFs than
    the cache holds (32), a CDF that changes between two calls (frequency EMA update), and concurrent callers all decode
    what was encoded; the automatic thread count (0) and explicit counts produce the same streams"""
    import threading
    rng = np.random.default_rng(7)
    cases = []
    for i in range(40):
        k = int(rng.choice([17, 512, 2048]))
        pmf = rng.random(k) ** 3 + 1e-4
        cdf = entropy.pmf_to_quantized_cdf(pmf / pmf.sum())[None]
        codes = torch.from_numpy(rng.integers(0, k, size=(3, 1, 5, 7)))
        streams = entropy.encode_level(codes, cdf)
        assert streams == entropy.encode_level(codes, cdf, threads=1) == entropy.encode_level(codes, cdf, threads=3)
        cases.append((codes, cdf, streams))
        assert torch.equal(entropy.decode_level(streams, 1, 5, 7, cdf), codes)
    for codes, cdf, streams in cases:                      # the first eight have been evicted by now
        assert torch.equal(entropy.decode_level(streams, 1, 5, 7, cdf, threads=2), codes)
    # same buffer, new contents: must not hit the old table
    codes, cdf, _ = cases[0]
    k = cdf.shape[1] - 1
    pmf = np.linspace(1.0, 2.0, k)
    cdf[0] = entropy.pmf_to_quantized_cdf(pmf / pmf.sum())
    assert torch.equal(entropy.decode_level(entropy.encode_level(codes, cdf), 1, 5, 7, cdf), codes)
    # large level: the automatic mode starts threads; the result is the single-threaded one
    big = torch.from_numpy(rng.integers(0, k, size=(6, 1, 96, 96)))
    s0 = entropy.encode_level(big, cdf)
    assert s0 == entropy.encode_level(big, cdf, threads=1)
    assert torch.equal(entropy.decode_level(s0, 1, 96, 96, cdf), big)
    errors = []

    def worker(case):
        try:
            for _ in range(20):
                assert torch.equal(entropy.decode_level(case[2], 1, 5, 7, case[1]), case[0])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(cases[i],)) for i in range(1, 9)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors
