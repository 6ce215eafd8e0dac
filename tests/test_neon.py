"""`Neon` tokenizer + `ResidualBackwardQuantizer` (mcquic/modules/compressor.py:181-233,
mcquic/modules/quantizer.py:577-765) and the codebook maintenance of SURVEY 8f NEXT-4 (quantizer.py:111-142).

CPU: state_dict layout == the reference's; the oracle restatement == the reference (bit for bit, where /root/reference
exists) == the committed reference outputs tests/golden/neon_*.npz (oracle/gen_golden.py --neon); the product's host
logic against the CPU model of the C ABI; reAssignCodebook against the reference method under the same RNG seed;
syncCodebook over a 2-rank gloo group.  GPU: encode/decode through the C ABI against the golden vectors.
"""
import os

import numpy as np
import pytest
import torch

from common import GOLDEN, NEON_CASES, neon_inputs
from emulator import EmulatedLib
from mcquic_b200 import Neon, _lib
from mcquic_b200.engine import Engine
from mcquic_b200.modules.quantizer import _multiCodebookQuantization
from oracle import mcquic_oracle as O
from oracle import ref_import

PIXEL_TOL = 2e-5        # relative to the output range: Neon decodes with the fp32-grade 3-pass path by default
ONE_PASS_TOL = 5e-3     # opt-in 1-pass (fp16 operands, TF32-grade) decode: 57 conv layers and 50 GroupNorms deep, the
                        # operand rounding accumulates to ~2.5e-3 of the output range (measured on the CPU model)
MARGIN_TIE = 2e-6       # a code flip at a relative top-2 distance gap below this is fp32 rounding, not a bug


def _golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    levels = len(NEON_CASES[name][2])
    codes = [torch.from_numpy(g[f"codes_{j}"].astype(np.int64)) for j in range(levels)]
    margins = [torch.from_numpy(g[f"margin_{j}"]) for j in range(levels)]
    return g, codes, margins


def _flips(codes, ref, margins):
    at = []
    for a, b, mg in zip(codes, ref, margins):
        assert a.dtype == torch.int64 and tuple(a.shape) == tuple(b.shape)
        at += mg[a.cpu() != b].tolist()
    return at


@pytest.mark.parametrize("name", list(NEON_CASES))
def test_oracle_reproduces_reference_golden(name):
    size = NEON_CASES[name][2]
    model, x = neon_inputs(name, Neon)
    sd = model.state_dict()
    g, ref, margins = _golden(name)
    codes = O.neon_encode(sd, x, size)
    assert _flips(codes, ref, margins) == []
    xhat = O.neon_decode(sd, ref, size)
    assert float((xhat - torch.from_numpy(g["xhat"])).abs().max()) <= 2e-6


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", list(NEON_CASES))
def test_oracle_and_state_dict_against_the_reference_itself(name):
    ref_import.load()
    from mcquic.modules.compressor import Neon as RefNeon
    size = NEON_CASES[name][2]
    ref_model, x = neon_inputs(name, RefNeon)
    mine, _ = neon_inputs(name, Neon)
    rsd, msd = ref_model.state_dict(), mine.state_dict()
    assert list(rsd) == list(msd)                                          # same keys, same order
    assert all(rsd[k].shape == msd[k].shape and rsd[k].dtype == msd[k].dtype for k in rsd)
    assert all(torch.equal(rsd[k], msd[k]) for k in rsd)                   # same deterministic weights
    with torch.inference_mode():
        codes = ref_model.encode(x)
        xhat = ref_model.decode(codes)
    assert all(torch.equal(a, b) for a, b in zip(codes, O.neon_encode(rsd, x, size)))
    assert torch.equal(xhat, O.neon_decode(rsd, codes, size))
    g, gcodes, _ = _golden(name)
    assert all(torch.equal(a, b) for a, b in zip(codes, gcodes))
    assert torch.equal(xhat, torch.from_numpy(g["xhat"]))


@pytest.mark.parametrize("name", list(NEON_CASES))
def test_host_logic_through_emulated_abi(name):
    c, k, size, dense, n, h, w = NEON_CASES[name]
    model, x = neon_inputs(name, Neon)
    model._engine = Engine(lib=EmulatedLib())
    g, ref, margins = _golden(name)
    hist = torch.zeros(len(size) * k, dtype=torch.int32)
    codes = model.encode(x, hist=hist)
    at = _flips(codes, ref, margins)
    assert at == [] or max(at) < MARGIN_TIE, at
    assert all(cd.is_contiguous() for cd in codes)
    # histogram segment j counts codes[j] (the order of the entropy coder's _freqEMA, quantizer.py:616)
    exp = torch.cat([torch.bincount(cd.flatten(), minlength=k) for cd in codes]).int()
    assert torch.equal(hist, exp)
    xref = torch.from_numpy(g["xhat"])
    scale = max(1.0, float(xref.abs().max()))
    assert model.decode_passes == 3
    assert float((model.decode(ref) - xref).abs().max()) <= PIXEL_TOL * scale
    model.decode_passes = 1
    assert float((model.decode(ref) - xref).abs().max()) <= ONE_PASS_TOL * scale
    with pytest.raises(RuntimeError):
        model.decode(ref[:-1])
    with pytest.raises(NotImplementedError):
        model.compress(x)                                # VariousMCoder.compress raises upstream too
    # the stage-2 generators' entry points (compressor.py:235-241), smallest level = level 0 of residual_forward
    from mcquic_b200 import engine as E
    old, E._DEFAULT = E._DEFAULT, model.engine
    try:
        model.engine.passes = 3
        sd = model.state_dict()
        lv = len(size) - 1
        first = model.residual_forward(ref[0], None, 0)
        want = O._neon_up(sd, f"_quantizer._decoders.{lv}", O.vq_dequantize(ref[0], sd["_quantizer._dequantizers.0._codebook"]),
                          O._neon_strided(size)[lv])
        assert float((first - want).abs().max()) <= 2e-5
        second = model.residual_forward(ref[1], first, 1)
        assert tuple(second.shape[:2]) == (n, 8)
        with pytest.raises(RuntimeError):
            model.residual_forward(ref[1], None, 1)
        with pytest.raises(RuntimeError):
            model.residual_forward(ref[0], first, 0)
    finally:
        E._DEFAULT = old


def test_size_sequence_is_validated_like_upstream():
    with pytest.raises(ValueError, match="does not half or equal"):
        Neon(32, 16, [16, 4], True)


# ------------------------------------------------------------------------------------------------ NEXT-4
def test_reassign_codebook_few_unused():
    torch.manual_seed(3)
    cb = torch.nn.Parameter(torch.randn(2, 8, 4))
    q = _multiCodebookQuantization(cb)
    before = cb.detach().clone()
    freq = torch.tensor([[.3, .0, .2, .1, .05, .0, .25, .1], [.125] * 8])
    moved = q.reAssignCodebook(freq)
    # codebook 0: slots 1 and 5 were never used -> overwritten by the two most used codewords (0, then 6)
    assert torch.equal(cb[0, 1], before[0, 0]) and torch.equal(cb[0, 5], before[0, 6])
    keep = [0, 2, 3, 4, 6, 7]
    assert torch.equal(cb[0, keep], before[0, keep]) and torch.equal(cb[1], before[1])
    assert moved.dtype == torch.bool and moved.shape == (16,) and moved.nonzero().flatten().tolist() == [1, 5]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
@pytest.mark.parametrize("unused", [2, 11])
def test_reassign_codebook_matches_reference_method(unused):
    """same codebook, same frequencies, same torch RNG seed -> identical codebook and mask, including the branch that
    spares a random half when more than k/2 codewords were never used (quantizer.py:118-125)"""
    ref_import.load()
    from mcquic.modules.quantizer import _multiCodebookQuantization as RefQ
    torch.manual_seed(11)
    init = torch.randn(3, 16, 8)
    freq = torch.rand(3, 16)
    freq[:, :unused] = 0
    freq[1] = freq[1][torch.randperm(16)]
    freq = freq / freq.sum(-1, keepdim=True)
    mine = _multiCodebookQuantization(torch.nn.Parameter(init.clone()))
    ref = RefQ(torch.nn.Parameter(init.clone()), None)
    torch.manual_seed(5)
    a = mine.reAssignCodebook(freq.clone())
    torch.manual_seed(5)
    b = ref.reAssignCodebook(freq.clone())
    assert torch.equal(a, b) and torch.equal(mine._codebook.data, ref._codebook.data)
    assert bool(a.any())


def _sync_worker(rank, world, port, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                       # ranks start from different codebooks
        model = Neon(32, 16, [8, 8], False)
        model.syncCodebook()
        out[rank] = [cb.detach().clone() for cb in model.Codebooks]
        freq = [f.clone() for f in model.NormalizedFreq]
        assert float(model.CodeUsage) == 1.0 and all(abs(float(f.sum()) - 1.0) < 1e-5 for f in freq)
    finally:
        dist.destroy_process_group()


def test_sync_codebook_two_ranks_gloo():
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 1000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sync_worker, args=(2, port, out), nprocs=2, join=True)
        a, b = out[0], out[1]
    assert len(a) == 2 and all(torch.equal(x, y) for x, y in zip(a, b))


# ------------------------------------------------------------------------------------------------ forward (values)
def _one_rank_group():
    import socket
    import torch.distributed as dist
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    return dist


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_quantizer_forward_values_match_the_reference_under_the_same_seed():
    """ResidualBackwardQuantizer.forward (quantizer.py:727-765; the soft path of SURVEY 8a row a11 in its only runnable
    owner): same weights, same input, same torch RNG seed -> same codes, logits, restored latent and frequency EMA as
    the reference module (forward values; the product path runs on the CPU model of the C ABI here)."""
    ref_import.load()
    from mcquic.modules.quantizer import ResidualBackwardQuantizer as RefQ
    from mcquic_b200 import ResidualBackwardQuantizer, engine as E
    from mcquic_b200.utils.synthetic import synthetic_block_state, uniform
    size = [8, 8, 4, 4]
    torch.manual_seed(0)
    ref = RefQ(32, size, True).eval()
    sd = synthetic_block_state(ref.state_dict(), "rbq.forward", seed=0)
    # a non-uniform frequency table so that _randomDrop really masks something
    for key in sd:
        if key.startswith("_entropyCoder._freqEMA"):
            f = uniform(tuple(sd[key].shape), key, 3).abs() ** 4
            sd[key] = f / f.sum(-1, keepdim=True)
    for key in list(sd):
        if key.endswith("._freqEMA") and key.startswith("_quantizers."):
            lv = int(key.split(".")[1])
            sd[key] = sd[f"_entropyCoder._freqEMA.{len(size) - 1 - lv}"]
    ref.load_state_dict(sd)
    mine = ResidualBackwardQuantizer(32, size, True).eval()
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.load_state_dict(sd)
    x = uniform((2, 8, 16, 16), "rbq.forward.x", 1) * 0.5
    dist = _one_rank_group()
    old, E._DEFAULT = E._DEFAULT, Engine(lib=EmulatedLib())
    try:
        E._DEFAULT.passes = 3
        torch.manual_seed(123)
        with torch.no_grad():
            r_y, r_codes, r_logits = ref(x)
        torch.manual_seed(123)
        y, codes, logits = mine(x)
    finally:
        E._DEFAULT = old
        dist.destroy_process_group()
    assert [tuple(c.shape) for c in codes] == [tuple(c.shape) for c in r_codes] == [(2, 1, 4, 4)] * 2 + [(2, 1, 8, 8)] * 2
    assert all(torch.equal(a, b) for a, b in zip(codes, r_codes))
    for a, b in zip(logits, r_logits):
        keep = b > -1e8                                     # entries _randomDrop pushed to -1e9 must coincide
        assert torch.equal(a > -1e8, keep) and int((~keep).sum()) > 0
        assert float((a[keep] - b[keep]).abs().max()) <= 2e-5 * float(b[keep].abs().max())
    assert float((y - r_y).abs().max()) <= 2e-5 * max(1.0, float(r_y.abs().max()))
    for a, b in zip(mine._entropyCoder._freqEMA, ref._entropyCoder._freqEMA):
        assert torch.allclose(a, b, atol=1e-7)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_neon_forward_values_match_the_reference_under_the_same_seed():
    """BaseCompressor.forward (compressor.py:35-43) on Neon: (xHat, yHat, codes, logits) equal the reference model's in
    training mode, same weights / input / RNG seed (forward values; product path on the CPU model of the C ABI)."""
    ref_import.load()
    from mcquic.modules.compressor import Neon as RefNeon
    name = "neon_c32_gn"
    ref_model, _ = neon_inputs(name, RefNeon)
    mine, _ = neon_inputs(name, Neon)
    from mcquic_b200.utils.synthetic import uniform
    x = uniform((2, 3, 64, 64), "neon.forward.image", 4)
    mine._engine = Engine(lib=EmulatedLib())
    dist = _one_rank_group()
    try:
        ref_model.train()
        torch.manual_seed(9)
        r_xhat, r_yhat, r_codes, r_logits = ref_model(x.clone())
        torch.manual_seed(9)
        xhat, yhat, codes, logits = mine(x)
    finally:
        dist.destroy_process_group()
    assert all(torch.equal(a, b) for a, b in zip(codes, r_codes))
    assert float((yhat - r_yhat.detach()).abs().max()) <= 2e-5 * max(1.0, float(r_yhat.detach().abs().max()))
    assert tuple(xhat.shape) == tuple(r_xhat.shape) == (2, 3, 64, 64)
    assert float((xhat - r_xhat.detach()).abs().max()) <= 2e-5 * max(1.0, float(r_xhat.detach().abs().max()))
    assert all(float((a - b.detach()).abs()[b > -1e8].max()) <= 2e-5 * float(b[b > -1e8].abs().max()) for a, b in zip(logits, [l.detach() for l in r_logits]))


def test_quantizer_forward_values_properties():
    """UMGMQuantizer.forward (quantizer.py:443-467) through the CPU model of the C ABI: shapes, code = argmax logit = the
    hard code of encode(), relaxed sample is one-hot up to an ulp, EMA moves toward the observed code frequencies."""
    from mcquic_b200 import Compressor, engine as E
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    model = Compressor(32, 2, [16, 8]).eval()
    model.load_state_dict(synthetic_state_dict(32, 2, [16, 8], seed=0))
    q = model._quantizer
    y = uniform((2, 32, 16, 16), "umgm.forward.y", 1) * 0.3
    old, E._DEFAULT = E._DEFAULT, Engine(lib=EmulatedLib())
    try:
        E._DEFAULT.passes = 3
        before = [f.clone() for f in q._entropyCoder._freqEMA]
        torch.manual_seed(7)
        yhat, codes, logits = q(y)
        hard = q.encode(y)
    finally:
        E._DEFAULT = old
    assert tuple(yhat.shape) == (2, 32, 16, 16)
    assert [tuple(c.shape) for c in codes] == [(2, 2, 8, 8), (2, 2, 4, 4)]
    assert [tuple(l.shape) for l in logits] == [(2, 2, 8, 8, 16), (2, 2, 4, 4, 8)]
    assert torch.equal(codes[0], logits[0].argmax(-1)) and torch.equal(codes[0], hard[0])   # level 0 sees the same input
    for lv, (f0, f1) in enumerate(zip(before, q._entropyCoder._freqEMA)):
        cnt = torch.stack([torch.bincount(codes[lv][:, j].flatten(), minlength=f0.shape[1]) for j in range(2)]).float()
        assert torch.allclose(f1, 0.1 * cnt / cnt.sum(-1, keepdim=True) + 0.9 * f0, atol=1e-6)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(NEON_CASES))
@pytest.mark.parametrize("graphs", [False, True])
def test_gpu_neon_against_reference_golden(name, graphs):
    c, k, size, dense, n, h, w = NEON_CASES[name]
    model, x = neon_inputs(name, Neon)
    model = model.cuda()
    model.use_graphs = graphs
    g, ref, margins = _golden(name)
    before = _lib.launch_count() + model.graph_launches
    hist = torch.zeros(len(size) * k, dtype=torch.int32, device="cuda")
    codes = model.encode(x.cuda(), hist=hist)
    if graphs:
        codes2 = model.encode(x.cuda())                      # pure replay
        assert all(torch.equal(a, b) for a, b in zip(codes, codes2))
    at = _flips(codes, ref, margins)
    assert at == [] or max(at) < MARGIN_TIE, f"code indices differ from the reference's at margins {at}"
    assert all(cd.is_cuda and cd.is_contiguous() for cd in codes)
    exp = torch.cat([torch.bincount(cd.flatten().cpu(), minlength=k) for cd in codes]).int()
    assert torch.equal(hist.cpu(), exp)
    xref = torch.from_numpy(g["xhat"])
    scale = max(1.0, float(xref.abs().max()))
    xhat = model.decode([cd.cuda() for cd in ref])
    assert tuple(xhat.shape) == tuple(xref.shape)
    assert float((xhat.cpu() - xref).abs().max()) <= PIXEL_TOL * scale
    model.decode_passes = 1
    assert float((model.decode([cd.cuda() for cd in ref]).cpu() - xref).abs().max()) <= ONE_PASS_TOL * scale
    assert _lib.launch_count() + model.graph_launches - before >= 100         # the CUDA path really ran
    assert model.engine.lib.mcq_device_error_flag() == 0
    with pytest.raises(RuntimeError):
        model.encode(x)                                      # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [False, True])
def test_gpu_neon_c128_groupnorm_statistics_fused_in_the_trunk(graphs):
    """Neon(128, ..., denseNorm=True): the trunk's 128 / 256-channel convs take the tcgen05 CTA-pair kernel, whose
    epilogue emits the GroupNorm partials (32 groups -> 4 / 8 channels per group) -- eager and under CUDA graphs,
    against the CPU oracle evaluated here (no golden needed at this size: ~1 s on CPU)."""
    from mcquic_b200.utils.synthetic import synthetic_block_state, uniform
    size = [8, 8]
    model = Neon(128, 64, size, True).eval()
    model.load_state_dict(synthetic_block_state(model.state_dict(), "neon.c128", seed=0))
    x = uniform((2, 3, 128, 128), "neon.c128.image", 1)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref, margins = O.neon_encode(sd, x, size, with_margin=True)
    xref = O.neon_decode(sd, ref, size)
    model = model.cuda()
    model.use_graphs = graphs
    # the fused-statistics path is really the one that runs for the trunk's first conv of every ResidualBlock
    probe = _lib.ConvParams()
    probe.n, probe.hin, probe.win, probe.cin, probe.cout, probe.cout_pad = 2, 128, 128, 128, 128, 128
    probe.ksize, probe.stride, probe.passes, probe.gn_groups = 3, 1, 3, 32
    probe.out_f32 = 1
    import ctypes
    rb, unit = ctypes.c_int32(0), ctypes.c_int32(0)
    assert _lib.load().mcq_conv_gn_layout(ctypes.byref(probe), ctypes.byref(rb), ctypes.byref(unit)) == 0
    assert (rb.value, unit.value) == (32 * 16, 4)
    for _ in range(2):
        codes = model.encode(x.cuda())
        at = _flips(codes, ref, margins)
        assert at == [] or max(at) < MARGIN_TIE, at
        xhat = model.decode([cd.cuda() for cd in ref])
        assert float((xhat.cpu() - xref).abs().max()) <= PIXEL_TOL * max(1.0, float(xref.abs().max()))
    assert model.engine.lib.mcq_device_error_flag() == 0


@pytest.mark.gpu
def test_gpu_quantizer_forward_values():
    """the values-only training-time forward on the GPU (CUDA RNG, so only RNG-independent facts are asserted): shapes,
    code = argmax of the returned logits, the un-dropped logits equal the deterministic ones, restored latent finite,
    frequency EMA still a distribution; UMGMQuantizer.forward level 0 reproduces encode()'s codes."""
    from mcquic_b200 import Compressor, ResidualBackwardQuantizer
    from mcquic_b200.utils.synthetic import synthetic_block_state, synthetic_state_dict, uniform
    size = [8, 8, 4, 4]
    q = ResidualBackwardQuantizer(32, size, True).eval()
    q.load_state_dict(synthetic_block_state(q.state_dict(), "rbq.forward", seed=0))
    q = q.cuda()
    x = (uniform((2, 8, 16, 16), "rbq.forward.x", 1) * 0.5).cuda()
    before = _lib.launch_count()
    y, codes, logits = q(x)
    torch.cuda.synchronize()
    assert _lib.launch_count() - before > 100 and _lib.load().mcq_device_error_flag() == 0
    assert tuple(y.shape) == (2, 8, 16, 16) and bool(torch.isfinite(y).all())
    assert [tuple(c.shape) for c in codes] == [(2, 1, 4, 4)] * 2 + [(2, 1, 8, 8)] * 2
    assert all(torch.equal(c, l.argmax(-1)) for c, l in zip(codes, logits))
    assert all(abs(float(f.sum()) - 1.0) < 1e-5 for f in q._entropyCoder._freqEMA)
    model = Compressor(32, 2, [16, 8]).eval()
    model.load_state_dict(synthetic_state_dict(32, 2, [16, 8], seed=0))
    model = model.cuda()
    yl = (uniform((2, 32, 16, 16), "umgm.forward.y", 1) * 0.3).cuda()
    yhat, codes, logits = model._quantizer(yl)
    assert tuple(yhat.shape) == (2, 32, 16, 16) and bool(torch.isfinite(yhat).all())
    assert torch.equal(codes[0], model._quantizer.encode(yl)[0])
    # BaseCompressor.forward (compressor.py:35-43), values only
    img = uniform((2, 3, 64, 96), "forward.image", 2).cuda()
    xhat, yhat, codes, logits = model(img)
    assert tuple(xhat.shape) == (2, 3, 64, 96) and tuple(yhat.shape) == (2, 32, 8, 12) and bool(torch.isfinite(xhat).all())
    assert [tuple(c.shape) for c in codes] == [(2, 2, 4, 6), (2, 2, 2, 3)]
    neon, ximg = neon_inputs("neon_c32_gn", Neon)
    neon = neon.cuda()
    xhat, yhat, codes, logits = neon(ximg[:, :, :96].cuda())
    assert tuple(xhat.shape) == (2, 3, 96, 128) and tuple(yhat.shape) == (2, 8, 12, 16) and bool(torch.isfinite(xhat).all())
    assert len(codes) == 4 and all(torch.equal(c, l.argmax(-1)) for c, l in zip(codes, logits))
    with pytest.raises(RuntimeError):
        neon(ximg[:, :, :90].cuda())                        # sides must be multiples of 16


@pytest.mark.gpu
def test_gpu_add_scaled():
    eng = Engine("tcgen05")
    eng.passes = 3
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(3, 5, 7, 8, generator=g).cuda(), torch.randn(3, 5, 7, 8, generator=g).cuda()
    out = eng.add_scaled(x, y, -1.0, (3, 5, 7, 8), {"f32", "raw"})
    assert torch.equal(out.f32, x - y)                       # one rounding: exactly the reference's subtraction
    rec = out.raw[0].double() + out.raw[1].double() / 2048.0
    assert float((rec - (x - y).double()).abs().max()) <= 2.0 ** -19 * float((x - y).abs().max())
    assert torch.equal(eng.add_scaled(x, y, 1.0, (3, 5, 7, 8), {"f32"}).f32, x + y)
    with pytest.raises(RuntimeError):
        eng.add_scaled(x, y[:2], 1.0, (3, 5, 7, 8), {"f32"})
