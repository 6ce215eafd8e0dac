"""VQ kernels through the C ABI vs the oracle (bit-exact indices; ties -> first index; logits to 1e-5)."""
import numpy as np
import pytest
import torch

from common import GOLDEN, log_parity
from mcquic_b200.engine import Engine
from mcquic_b200.utils.synthetic import uniform
from oracle import mcquic_oracle as O

pytestmark = pytest.mark.gpu


def _strict(what, shape, codes, ref, x, cb):
    """bit-exact indices: 0 flips, reported to the parity log (printed at the end of the run)"""
    mism = codes.cpu() != ref
    flips = int(mism.sum())
    if cb.shape[1] < 2:                      # a single codeword: no second-best distance
        log_parity(f"oracle vq {what} {shape}", flips, ref.numel())
        assert flips == 0
        return
    marg = O.vq_margin(x, cb)
    log_parity(f"oracle vq {what} {shape}", flips, ref.numel(), marg[mism].tolist(), float(marg.min()))
    assert flips == 0, (flips, marg[mism].tolist()[:8])


def _run(x_nchw, cb, logits=False):
    eng = Engine()
    n, c, h, w = x_nchw.shape
    xg = eng.from_nchw(x_nchw.cuda(), {"f32"}).f32
    cbg = cb.cuda().contiguous()
    c2 = (cbg ** 2).sum(-1).contiguous()
    m, k, _ = cb.shape
    hist = torch.zeros(m * k, dtype=torch.int32, device="cuda")
    out = eng.vq_assign(xg, cbg, c2, n, h, w, logits=logits, hist=hist)
    return out, hist, eng


@pytest.mark.parametrize("n,h,w,m,k,d", [
    (2, 16, 16, 1, 8192, 128),   # qp=1 level 0
    (2, 8, 8, 1, 2048, 128), (3, 4, 4, 1, 512, 128),
    (2, 32, 32, 6, 2048, 32),    # BASELINE configs[2] shape
    (1, 5, 7, 3, 100, 8),        # k not a multiple of the tile, ragged point count
    (1, 1, 1, 2, 1, 4),          # single codeword
    (2, 3, 3, 12, 64, 16),
])
def test_assign_matches_oracle(n, h, w, m, k, d):
    x = uniform((n, m * d, h, w), "vq.x", 11) * 0.26
    cb = uniform((m, k, d), "vq.cb", 11) * 0.19
    (codes, logit), hist, eng = _run(x, cb, logits=True)
    ref = O.vq_assign(x, cb)
    _strict("assign", (n, h, w, m, k, d), codes, ref, x, cb)
    assert codes.dtype == torch.int64 and codes.shape == (n, m, h, w)
    lref = O.vq_logits(x, cb, torch.ones(m, 1, 1, 1))
    assert float((logit.cpu() - lref).abs().max()) <= 1e-5 * max(1.0, float(lref.abs().max()))
    # histogram fused into the assign launch
    exp = torch.cat([h_.flatten() for h_ in O.code_histogram([codes.cpu()], [k])]).int()
    assert torch.equal(hist.cpu(), exp)
    # dequant: exact gather
    deq = eng.to_nchw(eng.vq_dequant(codes, cb.cuda().contiguous(), {"f32"}))
    assert torch.equal(deq.cpu(), O.vq_dequantize(codes.cpu(), cb))


@pytest.mark.parametrize("n,h,w,m,k,d", [
    (2, 16, 16, 1, 8192, 128),   # qp=1 level 0
    (64, 4, 4, 1, 512, 128),     # qp=1 level 2 at the benchmark batch (N tile narrowed to fill the SMs)
    (2, 5, 7, 2, 2048, 64),      # qp=2-like: two codebooks, channel slices of the same latent (35 points/image: not fused)
    (1, 5, 7, 1, 96, 64),        # ragged point count, k = 96 (N tile 96)
])
def test_tensor_core_assign_matches_oracle(n, h, w, m, k, d):
    """mcq_vq_assign_tc: x.c_k on tcgen05 (3-pass split fp16), distance + argmin in the GEMM epilogue (the path d = 128 took
    before the one-launch kernel covered it; still the route for d % 64 == 0 shapes the fused kernel does not tile)."""
    from mcquic_b200 import _lib
    from mcquic_b200.engine import split_weight
    _lib.set_option("vq128_fused", 0)
    x = uniform((n, m * d, h, w), "vqtc.x", 7) * 0.26
    cb = uniform((m, k, d), "vqtc.cb", 7) * 0.19
    eng = Engine()
    xg = eng.from_nchw(x.cuda(), {"f32"}).f32
    cbg = cb.cuda().contiguous()
    c2 = (cbg ** 2).sum(-1).contiguous()
    hist = torch.zeros(m * k, dtype=torch.int32, device="cuda")
    packed = split_weight(cbg.reshape(-1, d))
    before = eng.lib.mcq_kernel_launch_count()
    codes = eng.vq_assign(xg, cbg, c2, n, h, w, hist=hist, packed=packed)
    assert eng.lib.mcq_kernel_launch_count() - before == 2 + m      # prep + m GEMMs + finalize: the tcgen05 path ran
    ref = O.vq_assign(x, cb)
    _strict("tensor-core", (n, h, w, m, k, d), codes, ref, x, cb)
    simt = eng.vq_assign(xg, cbg, c2, n, h, w)
    assert torch.equal(codes, simt)
    exp = torch.cat([h_.flatten() for h_ in O.code_histogram([codes.cpu()], [k])]).int()
    assert torch.equal(hist.cpu(), exp)
    assert eng.lib.mcq_device_error_flag() == 0
    _lib.set_option("vq128_fused", 1)


@pytest.mark.parametrize("n,h,w,m,k,d", [
    (2, 16, 16, 1, 8192, 128),   # qp=1 level 0 (d = 128: single latent buffer, codebook in half-row stages)
    (64, 4, 4, 1, 512, 128),     # qp=1 level 2 at the benchmark batch
    (3, 8, 8, 2, 2048, 128),     # two codebooks of d = 128
    (2, 32, 32, 6, 2048, 32),    # BASELINE configs[2] shape (qp=3-like: M=6, K=2048, d=32), 2 images
    (2, 8, 8, 2, 2048, 64),      # qp=2-like: d=64 (two 64-element K chunks per operand row)
    (3, 4, 4, 6, 256, 32),       # 16-point images: a 32-row store box spans two images; ragged last tile (48 points)
    (5, 16, 16, 3, 128, 32),     # single codeword chunk per tile (k = 128)
    (1, 1, 1, 2, 384, 64),       # one point
    (40, 16, 16, 6, 512, 32),    # 480 tiles: several tiles per persistent CTA
])
@pytest.mark.parametrize("logits", [False, True])
def test_fused_assign_matches_oracle(n, h, w, m, k, d, logits):
    """mcq_vq_assign_fused (csrc/vq_fused.cuh): one launch = split + tcgen05 x.c + distance + argmin (+ logits via TMA
    store) + histogram.  Codes bit-exact (flips only at fp32 near-ties), logits to 1e-5, histogram exact."""
    from mcquic_b200.engine import pack_codebook
    x = uniform((n, m * d, h, w), "vqf.x", 13) * 0.26
    cb = uniform((m, k, d), "vqf.cb", 13) * 0.19
    temp = (uniform((m,), "vqf.t", 13).abs() + 0.5).cuda()
    eng = Engine()
    xg = eng.from_nchw(x.cuda(), {"f32"}).f32
    cbg = cb.cuda().contiguous()
    c2 = (cbg ** 2).sum(-1).contiguous()
    hist = torch.zeros(m * k, dtype=torch.int32, device="cuda")
    assert eng.lib.mcq_vq_fused_supported(h, w, k, d) == 1
    before = eng.lib.mcq_kernel_launch_count()
    out = eng.vq_assign(xg, cbg, c2, n, h, w, logits=logits, logit_scale=temp if logits else None, hist=hist,
                        packed=pack_codebook(cbg))
    torch.cuda.synchronize()
    assert eng.lib.mcq_kernel_launch_count() - before == 1
    assert eng.lib.mcq_device_error_flag() == 0
    codes = out[0] if logits else out
    ref = O.vq_assign(x, cb)
    _strict(f"fused logits={logits}", (n, h, w, m, k, d), codes, ref, x, cb)
    assert codes.dtype == torch.int64 and codes.shape == (n, m, h, w)
    if logits:
        lref = O.vq_logits(x, cb, temp.cpu().reshape(m, 1, 1, 1))
        assert out[1].shape == (n, m, h, w, k)
        assert float((out[1].cpu() - lref).abs().max()) <= 1e-5 * max(1.0, float(lref.abs().max()))
    exp = torch.cat([h_.flatten() for h_ in O.code_histogram([codes.cpu()], [k])]).int()
    assert torch.equal(hist.cpu(), exp)


def test_fused_ties_pick_the_first_index():
    m, k, d = 2, 256, 32
    cb = uniform((m, k, d), "tie.cb32", 2)
    cb[:, 200] = cb[:, 17]          # duplicates in the other column half / a later chunk: the first index must win
    cb[:, 150] = cb[:, 17]
    cb[:, 90] = cb[:, 17]
    x = cb[:, 17].reshape(1, m * d, 1, 1).repeat(4, 1, 4, 4).clone()
    eng = Engine()
    cbg = cb.cuda().contiguous()
    codes = eng.vq_assign(eng.from_nchw(x.cuda(), {"f32"}).f32, cbg, (cbg ** 2).sum(-1).contiguous(), 4, 4, 4)
    assert (codes == 17).all()


def test_tensor_core_ties_pick_the_first_index():
    from mcquic_b200.engine import split_weight
    m, k, d = 1, 256, 64
    cb = uniform((m, k, d), "tie.cb64", 2)
    cb[:, 200] = cb[:, 17]
    cb[:, 150] = cb[:, 17]
    x = cb[:, 17].reshape(1, m * d, 1, 1).repeat(3, 1, 2, 2).clone()
    eng = Engine()
    cbg = cb.cuda().contiguous()
    codes = eng.vq_assign(eng.from_nchw(x.cuda(), {"f32"}).f32, cbg, (cbg ** 2).sum(-1).contiguous(), 3, 2, 2,
                          packed=split_weight(cbg.reshape(-1, d)))
    assert (codes == 17).all()


def test_exact_ties_pick_the_first_index():
    m, k, d = 2, 300, 8
    cb = uniform((m, k, d), "tie.cb", 2)
    cb[:, 200] = cb[:, 17]          # duplicate codewords: distances tie exactly
    cb[:, 250] = cb[:, 17]
    x = cb[:, 17].reshape(1, m * d, 1, 1).repeat(3, 1, 2, 2).clone()
    (codes), _, _ = _run(x, cb)
    assert (codes == 17).all()
    assert torch.equal(codes.cpu(), O.vq_assign(x, cb))


def test_golden_vq_vector():
    g = np.load(f"{GOLDEN}/vq_m6_k2048_d32.npz")
    m, k, d, n, h, w = g["config"].tolist()
    cb = uniform((m, k, d), "vq.codebook", 3) * ((2.0 / (5 * d)) ** 0.5 * 3 ** 0.5)
    x = uniform((n, m * d, h, w), "vq.latent", 3) * 0.26
    codes, _, _ = _run(x, cb)
    ref = torch.from_numpy(g["codes"].astype(np.int64))
    mism = codes.cpu() != ref
    marg = torch.from_numpy(g["margin"])
    log_parity("golden vq_m6_k2048_d32 (4 images)", int(mism.sum()), ref.numel(), marg[mism].tolist(), float(marg.min()))
    assert int(mism.sum()) == 0


def test_sub_api_modules():
    from mcquic_b200.modules.quantizer import _multiCodebookDeQuantization, _multiCodebookQuantization
    cb = torch.nn.Parameter((uniform((3, 50, 8), "sub.cb", 1) * 0.2).cuda())
    q, dq = _multiCodebookQuantization(cb).cuda(), _multiCodebookDeQuantization(cb)
    x = (uniform((2, 24, 5, 7), "sub.x", 1) * 0.2).cuda()
    code = q.encode(x)
    assert torch.equal(code.cpu(), O.vq_assign(x.cpu(), cb.data.cpu()))
    assert torch.equal(dq.decode(code).cpu(), O.vq_dequantize(code.cpu(), cb.data.cpu()))
    with pytest.raises(RuntimeError):
        dq.decode(torch.full_like(code, 50))
    with pytest.raises(RuntimeError):
        q.encode(x[:, :20])


def test_full_size_against_reference_golden():
    """BASELINE configs[2] at FULL size (N=32, 32x32 grid, M=6, K=2048, d=32): all 196 608 indices equal the reference
    quantizer's (tests/golden/vq_cfg3_full.npz, `_multiCodebookQuantization.encode` of the imported reference), hard path
    and soft path (codes + logits in one launch); histogram == bincount; logits of 2 images vs the oracle to 1e-5."""
    g = np.load(f"{GOLDEN}/vq_cfg3_full.npz")
    m, k, d, n, h, w = g["config"].tolist()
    assert (n, h, w, m, k, d) == (32, 32, 32, 6, 2048, 32)
    cb = uniform((m, k, d), "vq.codebook", 3) * ((2.0 / (5 * d)) ** 0.5 * 3 ** 0.5)
    x = uniform((n, m * d, h, w), "vq.latent.full", 3) * 0.26
    ref = torch.from_numpy(g["codes"].astype(np.int64))
    marg = torch.from_numpy(g["margin"].astype(np.float32))
    for logits in (False, True):
        out, hist, eng = _run(x, cb, logits=logits)
        codes = out[0] if logits else out
        mism = codes.cpu() != ref
        log_parity(f"golden vq_cfg3_full (BASELINE configs[2], N=32) logits={logits}", int(mism.sum()), ref.numel(),
                   marg[mism].tolist(), float(g["min_margin"]))
        assert int(mism.sum()) == 0, marg[mism].tolist()[:8]
        exp = torch.cat([h_.flatten() for h_ in O.code_histogram([codes.cpu()], [k])]).int()
        assert torch.equal(hist.cpu(), exp) and int(hist.sum()) == n * h * w * m
        if logits:
            lref = O.vq_logits(x[:2], cb, torch.ones(m, 1, 1, 1))
            assert float((out[1][:2].cpu() - lref).abs().max()) <= 1e-5 * max(1.0, float(lref.abs().max()))
        del out
    assert eng.lib.mcq_device_error_flag() == 0


def test_full_size_vs_oracle_chunked_by_image():
    """a second full-size input (different seed), compared with the CPU oracle image by image (the [n, m, hw, k] distance
    tensor of the whole batch would be 1.6 GB on the host)"""
    n, h, w, m, k, d = 32, 32, 32, 6, 2048, 32
    x = uniform((n, m * d, h, w), "vq.big", 5) * 0.26
    cb = uniform((m, k, d), "vq.bigcb", 5) * 0.19
    codes, hist, eng = _run(x, cb)
    assert int(hist.sum()) == n * h * w * m
    ref = torch.cat([O.vq_assign(x[i:i + 1], cb) for i in range(n)])
    mism = codes.cpu() != ref
    flips = int(mism.sum())
    at = [] if flips == 0 else torch.cat([O.vq_margin(x[i:i + 1], cb) for i in range(n)])[mism].tolist()
    log_parity("oracle vq full size (N=32, M=6, K=2048, d=32), second input", flips, ref.numel(), at)
    assert flips == 0, at[:8]
