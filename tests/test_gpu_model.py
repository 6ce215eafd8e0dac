"""End-to-end parity of Compressor.encode/decode on the GPU against the committed golden vectors (reference
outputs) and the oracle: code indices bit-exact, pixels within 1e-3 absolute (north_star), plus size-independent
properties at the benchmark size."""
import pytest
import torch

import numpy as np

from common import code_report, golden_codes, golden_inputs, load_golden, log_parity
from mcquic_b200 import Compressor, _lib
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
from oracle import mcquic_oracle as O

pytestmark = pytest.mark.gpu

PIXEL_TOL = 1e-3   # BASELINE.json north_star: reconstructed pixels within 1e-3 abs fp32


def _model(cfg, sd, impl="tcgen05"):
    model = Compressor(cfg["channel"], cfg["m"], cfg["k"]).eval()
    model.load_state_dict(sd)
    model = model.cuda()
    model.set_impl(impl)
    return model


@pytest.mark.parametrize("name", ["compressor_small", "compressor_qp1_256", "compressor_c192_m6"])
@pytest.mark.parametrize("graphs", [False, True])
def test_against_reference_golden(name, graphs):
    g, cfg = load_golden(name)
    sd, x = golden_inputs(cfg)
    model = _model(cfg, sd)
    model.use_graphs = graphs
    before = _lib.launch_count() + model.graph_launches
    codes = model.encode(x.cuda())
    if graphs:
        codes = model.encode(x.cuda())                     # second call = pure graph replay
    ref = golden_codes(g, len(cfg["k"]))
    margins = [g[f"margin_{lv}"] for lv in range(len(ref))]
    flips, total, at = code_report(codes, ref, margins)
    log_parity(f"golden {name} graphs={graphs}", flips, total, at, min(float(mg.min()) for mg in margins))
    assert flips == 0, f"{flips}/{total} code indices differ from the reference; margins there: {at}"
    assert all(c.dtype == torch.int64 and c.is_cuda and c.is_contiguous() for c in codes)
    xhat = model.decode([c.cuda() for c in ref])
    s = cfg["stride"]
    err = float((xhat.cpu()[..., ::s, ::s] - torch.from_numpy(g["xhat_sample"])).abs().max())
    assert err <= PIXEL_TOL, err
    assert _lib.launch_count() + model.graph_launches - before >= 40      # the CUDA path really ran
    assert model.engine.lib.mcq_device_error_flag() == 0


def test_simt_and_tcgen05_agree_with_oracle_including_psnr():
    g, cfg = load_golden("compressor_qp1_256")
    sd, x = golden_inputs(cfg)
    ref = O.encode(sd, x)
    xref = O.decode(sd, ref)
    for impl in ("simt", "tcgen05"):
        model = _model(cfg, sd, impl)
        codes = model.encode(x.cuda())
        flips, total, _ = code_report(codes, ref)
        log_parity(f"oracle qp1 256x256 impl={impl}", flips, total)
        assert flips == 0
        for passes, tol in ((3, 5e-6), (1, PIXEL_TOL)):
            model.decode_passes = passes
            xhat = model.decode(codes).cpu()
            assert float((xhat - xref).abs().max()) <= tol
            # BASELINE configs[0]: PSNR of uint8 images (DeTransform + PSNR of the reference) >= 60 dB
            assert float(O.psnr_uint8(O.to_uint8(xhat), O.to_uint8(xref)).min()) >= 60.0


def test_unaligned_image_is_reflect_padded_like_the_reference():
    cfg = dict(channel=64, m=2, k=[256, 128, 64])
    sd = synthetic_state_dict(64, 2, [256, 128, 64], seed=0)
    x = uniform((1, 3, 150, 333), "pad.image", 4)
    model = _model(cfg, sd)
    codes = model.encode(x.cuda())
    ref, marg = O.encode(sd, x, with_margin=True)
    assert [tuple(c.shape) for c in codes] == [tuple(r.shape) for r in ref] == [(1, 2, 16, 24), (1, 2, 8, 12), (1, 2, 4, 6)]
    flips, total, at = code_report(codes, ref, marg)
    log_parity("oracle unaligned 150x333 (C=64, m=2)", flips, total, at, min(float(mg.min()) for mg in marg))
    assert flips == 0, (flips, at)
    assert float((model.decode(codes).cpu() - O.decode(sd, [c.cpu() for c in codes])).abs().max()) <= PIXEL_TOL


def test_large_unaligned_image_like_the_cli_sees():
    """one 1100 x 1900 photograph-sized image (the reference CLI's use case, demo.py:109-134): reflect-padded to
    1152 x 1920, 576 x 960 feature maps after the stem, ragged tiles on every level -- codes vs the CPU oracle,
    pixels within 1e-3"""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(128, 1, cfg["k"], seed=0)
    x = uniform((1, 3, 1100, 1900), "large.image", 7)
    model = _model(cfg, sd)
    codes = model.encode(x.cuda())
    ref, marg = O.encode(sd, x, with_margin=True)
    assert [tuple(c.shape) for c in codes] == [tuple(r.shape) for r in ref] == [(1, 1, 72, 120), (1, 1, 36, 60), (1, 1, 18, 30)]
    flips, total, at = code_report(codes, ref, marg)
    log_parity("oracle large unaligned 1100x1900 (qp=1)", flips, total, at, min(float(mg.min()) for mg in marg))
    assert flips == 0, (flips, at)
    xhat = model.decode([r.cuda() for r in ref])
    assert tuple(xhat.shape) == (1, 3, 1152, 1920)
    assert float((xhat.cpu() - O.decode(sd, ref)).abs().max()) <= PIXEL_TOL
    assert model.engine.lib.mcq_device_error_flag() == 0


def test_compress_decompress_and_identical_bpp():
    """compress()/decompress() on the CUDA path: the reference flow yields 424 + 96 + 24 bytes (0.0664 bpp) for the
    qp=1 golden image under the uniform prior; identical codes + bit-identical rANS => identical bpp."""
    from mcquic_b200 import entropy
    g, cfg = load_golden("compressor_qp1_256")
    sd, x = golden_inputs(cfg)
    model = _model(cfg, sd)
    codes, binaries, headers = model.compress(x.cuda())
    assert [len(b) for b in binaries[0]] == [424, 96, 24]
    assert entropy.bpp(binaries[0], headers[0].ImageSize) == pytest.approx(0.06640625)
    out = model.decompress(binaries, headers)
    assert torch.equal(out, model.decode(codes))                       # 256x256: no crop
    s = cfg["stride"]
    assert float((out.cpu()[..., ::s, ::s] - torch.from_numpy(g["xhat_sample"])).abs().max()) <= PIXEL_TOL


def test_errors():
    cfg = dict(channel=64, m=2, k=[256, 128, 64])
    model = _model(cfg, synthetic_state_dict(64, 2, [256, 128, 64], seed=0))
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(1, 1, 128, 128, device="cuda"))
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(1, 3, 128, 128))                            # CPU tensor: no fallback
    codes = model.encode(torch.zeros(1, 3, 128, 128, device="cuda"))
    with pytest.raises(RuntimeError):
        model.decode(codes[:2])
    bad = [c.clone() for c in codes]
    bad[1][0, 0, 0, 0] = 128
    with pytest.raises(RuntimeError):
        model.decode(bad)
    with pytest.raises(RuntimeError):
        model.decode([])


def test_benchmark_size_properties():
    """qp=1, batch 64 x 3 x 256 x 256 (BASELINE configs[1]): (a) batching invariance -- images are independent,
    so the first 2 images alone give the same codes as inside the batch of 64 (this is also what makes the
    multi-GPU sharding exact); (b) determinism across graph replays; (c) histogram == bincount of the codes;
    (d) decode(encode(x)) is a fixed function: decoding twice is bit-identical; (e) ALL 21 504 codes of the 64 images
    bench.py times equal the reference's (tests/golden/bench_qp1_n64.npz, made by oracle/gen_golden.py --baseline from the
    imported reference) with 0 flips, and the decoded pixels are within 1e-3 of the reference's."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(128, 1, cfg["k"], seed=0)
    model = _model(cfg, sd)
    x = uniform((64, 3, 256, 256), "bench.image.0", 0)
    hist = torch.zeros(sum(cfg["k"]), dtype=torch.int32, device="cuda")
    codes = model.encode(x.cuda(), hist=hist)
    codes2 = model.encode(x.cuda())
    assert all(torch.equal(a, b) for a, b in zip(codes, codes2))
    small = model.encode(x[:2].cuda())
    assert all(torch.equal(a[:2], b) for a, b in zip(codes, small))
    exp = torch.cat([h.flatten() for h in O.code_histogram([c.cpu() for c in codes], cfg["k"])]).int()
    assert torch.equal(hist.cpu(), exp) and int(hist.sum()) == 64 * (256 + 64 + 16)
    xhat = model.decode(codes)
    assert torch.equal(xhat, model.decode(codes)) and tuple(xhat.shape) == (64, 3, 256, 256)
    g, gcfg = load_golden("bench_qp1_n64")
    assert (gcfg["channel"], gcfg["m"], gcfg["n"], gcfg["h"], gcfg["w"], gcfg["k"]) == (128, 1, 64, 256, 256, cfg["k"])
    ref = golden_codes(g, 3)
    flips, total, at = code_report(codes, ref, [g[f"margin_{lv}"] for lv in range(3)])
    log_parity("golden bench_qp1_n64 (BASELINE configs[1], all 64 images)", flips, total, at, float(g["min_margin"]))
    assert total == 64 * 336 and flips == 0, (flips, at)
    s = gcfg["stride"]
    err = float((xhat.cpu()[..., ::s, ::s] - torch.from_numpy(g["xhat_sample"])).abs().max())
    assert err <= PIXEL_TOL, err
    # the host-buffer pipeline (what bench.py's e2e leg calls) gives the same 21 504 codes
    codes_h = model.encode(x.pin_memory())
    assert all(torch.equal(a, b) for a, b in zip(codes, codes_h))


def test_q6_512_against_reference_golden():
    """Q6 = Compressor(192, 6, [2048] * 3), the model behind BASELINE configs[2], on 10 images of 512 x 512: all 80 640
    codes equal the reference's (tests/golden/compressor_q6_512.npz), pixels within 1e-3."""
    g, cfg = load_golden("compressor_q6_512")
    sd = synthetic_state_dict(cfg["channel"], cfg["m"], cfg["k"], seed=0)
    x = uniform((cfg["n"], 3, cfg["h"], cfg["w"]), "q6.image", 5)
    model = _model(cfg, sd)
    codes = model.encode(x.cuda())
    ref = golden_codes(g, 3)
    flips, total, at = code_report(codes, ref, [g[f"margin_{lv}"] for lv in range(3)])
    log_parity("golden compressor_q6_512 (C=192, M=6, K=2048, 10 x 512^2)", flips, total, at, float(g["min_margin"]))
    assert total == 10 * 6 * (1024 + 256 + 64) and flips == 0, (flips, at)
    xhat = model.decode([c.cuda() for c in ref])
    s = cfg["stride"]
    err = float((xhat.cpu()[..., ::s, ::s] - torch.from_numpy(g["xhat_sample"])).abs().max())
    assert err <= PIXEL_TOL, err
    assert model.engine.lib.mcq_device_error_flag() == 0


@pytest.mark.parametrize("graphs", [False, True])
@pytest.mark.parametrize("via", ["compressor", "quantizer"])
def test_reencode_after_codebook_reassignment(graphs, via):
    """NEXT-4 on the device: after reAssignCodebook() a CUDA-resident model must encode with the NEW codebook -- no stale
    packed codebook, no stale CUDA graph.  `via="quantizer"` calls the quantizer's method directly (what the reference's
    hook does through Compound, compound.py:52-58), bypassing BaseCompressor.reAssignCodebook's explicit invalidate():
    the weight fingerprint has to catch it."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(128, 1, cfg["k"], seed=0)
    model = _model(cfg, sd)
    model.use_graphs = graphs
    x = uniform((2, 3, 256, 256), "reassign.image", 3)
    before = model.encode(x.cuda())
    before = model.encode(x.cuda())                 # graph replay path warmed
    # frequencies: the codewords this batch used stay "used", most others are marked never-used -> they get overwritten
    for lv, code in enumerate(before):
        f = model._quantizer._entropyCoder._freqEMA[lv]
        with torch.no_grad():
            f.zero_()
            f[0, code.flatten()[::3]] = 1.0          # only a third of the used ones stay: the rest must move
            f[0, : f.shape[1] // 4] += 0.5
    moved = model.reAssignCodebook() if via == "compressor" else model._quantizer.reAssignCodebook()
    assert float(moved) > 0.1
    new_sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    assert not torch.equal(new_sd["_quantizer._encoders.0._quantizer._codebook"],
                           sd["_quantizer._encoders.0._quantizer._codebook"])
    codes = model.encode(x.cuda())
    ref, marg = O.encode(new_sd, x, with_margin=True)
    flips, total, at = code_report(codes, ref, marg)
    log_parity(f"oracle after reAssignCodebook via={via} graphs={graphs}", flips, total, at,
               min(float(mg.min()) for mg in marg))
    assert flips == 0, (flips, at)
    assert any(not torch.equal(a, b) for a, b in zip(codes, before))       # the reassignment really changed the result
    assert float((model.decode(codes).cpu() - O.decode(new_sd, ref)).abs().max()) <= PIXEL_TOL


@pytest.mark.parametrize("case", ["x*2^10", "x*2^-10", "stem*2^10,x*2^-10", "codebooks*2^2", "codebooks*2^-8"])
def test_fp16_range_stress(case):
    """Activations travel as split-fp16 planes without a per-tensor scale (include/mcquic_b200.h): check the path against
    the fp32 oracle where activations are 2^10 times larger / smaller than with [-1, 1] images and reference-scale weights
    (squares of the GDN operand would leave fp16's range without MCQ_SQUARE_SCALE), with pre-scaled weights, and on the
    decode side with codebooks 4 times larger (the IGDN stages amplify quadratically: pixels reach ~1e3; at 2^8 the
    reference's own fp32 output is 3.5e8 and the fp16 planes saturate -- activations beyond 6.5e4 are outside the format,
    include/mcquic_b200.h) and 2^-8 times smaller.  Codes: 0 flips; pixels: 1e-3 of the oracle's output range."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(128, 1, cfg["k"], seed=0)
    x = uniform((2, 3, 256, 256), "range.image", 9)
    if case == "x*2^10":
        x = x * 1024.0
    elif case == "x*2^-10":
        x = x / 1024.0
    elif case == "stem*2^10,x*2^-10":
        sd["_encoder.0.weight"] = sd["_encoder.0.weight"] * 1024.0
        x = x / 1024.0
    else:
        factor = 4.0 if case == "codebooks*2^2" else 2.0 ** -8
        for key in sd:
            if key.endswith("._codebook"):
                sd[key] = sd[key] * factor
    model = _model(cfg, sd)
    codes = model.encode(x.cuda())
    ref, marg = O.encode(sd, x, with_margin=True)
    flips, total, at = code_report(codes, ref, marg)
    log_parity(f"oracle fp16-range {case}", flips, total, at, min(float(mg.min()) for mg in marg))
    assert flips == 0, (flips, at)
    xref = O.decode(sd, ref)
    xhat = model.decode([c.cuda() for c in ref]).cpu()
    assert torch.isfinite(xhat).all()
    assert float((xhat - xref).abs().max()) <= PIXEL_TOL * max(1.0, float(xref.abs().max()))
    assert model.engine.lib.mcq_device_error_flag() == 0


@pytest.mark.parametrize("graphs", [False, True])
def test_batch_slices_of_the_full_resolution_layers_change_nothing(graphs):
    """`encode_slice` / `decode_slice` (L2-resident sub-batches of the full-resolution layers; off by default since the
    bulk-store drain) must give the very same codes and pixels as the whole batch at once, also for a ragged last slice."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(cfg["channel"], cfg["m"], cfg["k"], seed=0)
    model = _model(cfg, sd)
    model.use_graphs = graphs
    x = uniform((7, 3, 128, 192), "slices.image", 0).cuda()
    ref_codes = model.encode(x)
    ref_x = model.decode(ref_codes)
    for sl in (2, 4):
        model.encode_slice = model.decode_slice = sl
        codes = model.encode(x)
        assert all(torch.equal(a, b) for a, b in zip(codes, ref_codes)), sl
        assert torch.equal(model.decode(codes), ref_x), sl
    assert model.engine.lib.mcq_device_error_flag() == 0


@pytest.mark.parametrize("hw,n", [((128, 128), 16), ((100, 72), 6), ((64, 128), 32)])
def test_host_pipeline_matches_device_path(hw, n):
    """encode(pinned host batch) / decode(out=pinned host tensor): the chunked copy/compute pipeline (first and last
    full-resolution layers run per batch slice while PCIe moves the next / previous slice) must give the very same codes
    and pixels as the device-resident path, every step (replays reuse the staging buffers)."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(cfg["channel"], cfg["m"], cfg["k"], seed=0)
    model = _model(cfg, sd)
    assert model.host_slices(64, True) == [(0, 4), (4, 16), (16, 32), (32, 48), (48, 64)]
    assert model.host_slices(64, False) == [(0, 16), (16, 32), (32, 48), (48, 60), (60, 64)]
    assert model.host_slices(16, True) == [(0, 2), (2, 8), (8, 16)] and model.host_slices(6, True) == [(0, 6)]
    total = sum(cfg["m"] * k for k in cfg["k"])
    for step in range(3):
        x = uniform((n, 3) + hw, f"pipe.image.{step}", 0)
        xh = x.pin_memory()
        hist_d = torch.zeros(total, dtype=torch.int32, device="cuda")
        hist_h = torch.zeros(total, dtype=torch.int32, device="cuda")
        ref_codes = model.encode(x.cuda(), hist=hist_d)
        codes = model.encode(xh, hist=hist_h)
        assert all(c.is_cuda for c in codes)
        assert all(torch.equal(a, b) for a, b in zip(codes, ref_codes))
        assert torch.equal(hist_d, hist_h)
        ref_x = model.decode(ref_codes)
        out = torch.empty(tuple(ref_x.shape), dtype=torch.float32).pin_memory()
        got = model.decode(codes, out=out)
        assert got is out
        assert torch.equal(out, ref_x.cpu())
    assert model.engine.lib.mcq_device_error_flag() == 0
    with pytest.raises(RuntimeError):
        model.decode(codes, out=torch.empty(tuple(ref_x.shape)))           # not pinned
    with pytest.raises(RuntimeError):
        model.encode(uniform((2, 3) + hw, "pipe.unpinned", 0))             # plain CPU tensor: no CPU fallback


def test_uint8_images_in_and_out_like_the_reference_cli_flow():
    """demo.compressImage / decompressImage (demo.py:109-134) hold images as uint8: `convert_image_dtype` + `(x - 0.5) * 2`
    before `compress`, `DeTransform` after `decompress`.  encode(uint8) / decode(out=uint8) apply both inside the first / last
    kernel: the codes must equal those of the float path fed with the reference's transform of the same bytes (and the
    oracle's), the uint8 pixels must equal DeTransform of the float pixels exactly -- device-resident and through the host
    pipeline."""
    cfg = dict(channel=128, m=1, k=[8192, 2048, 512])
    sd = synthetic_state_dict(128, 1, cfg["k"], seed=0)
    model = _model(cfg, sd)
    for (n, h, w) in ((16, 256, 256), (3, 200, 328)):
        u8 = ((uniform((n, 3, h, w), "u8.image", 6) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
        xf = (u8.float() / 255.0 - 0.5) * 2
        ref_codes = model.encode(xf.cuda())
        for src in (u8.cuda(), u8.pin_memory()):
            codes = model.encode(src)
            assert all(torch.equal(a, b) for a, b in zip(codes, ref_codes))
        oc, marg = O.encode(sd, xf[:2], with_margin=True)
        flips, total, at = code_report([c[:2] for c in codes], oc, marg)
        log_parity(f"oracle uint8 input {n}x{h}x{w}", flips, total, at, min(float(mg.min()) for mg in marg))
        assert flips == 0, (flips, at)
        xhat = model.decode(codes)
        exp = O.to_uint8(xhat.cpu())
        out = torch.empty(tuple(xhat.shape), dtype=torch.uint8).pin_memory()
        got = model.decode(codes, out=out)
        assert got is out and torch.equal(out, exp)
        outf = torch.empty(tuple(xhat.shape), dtype=torch.float32).pin_memory()
        assert torch.equal(model.decode(codes, out=outf), xhat.cpu())
    assert model.engine.lib.mcq_device_error_flag() == 0
