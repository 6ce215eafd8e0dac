"""CPU checks of the host side: the C-ABI library loads and exports every symbol the header declares, the
Python mirror has the reference's state_dict layout, and the module walk / fusion map / weight repacking is
correct -- verified by running the product's Python against a CPU model of the C ABI (tests/emulator.py) and
comparing with the oracle.  No CUDA compute happens here."""
import ctypes
import os
import re

import pytest
import torch

from common import code_report, golden_codes, golden_inputs, load_golden
from emulator import EmulatedLib
from mcquic_b200 import Compressor, _lib
from mcquic_b200.engine import Engine, pack_conv, split_weight
from mcquic_b200.modules.compressor import aligned_pad_amounts
from oracle import mcquic_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mcquic_b200.h")).read()
    declared = set(re.findall(r"\b(mcq_[a-z0-9_]+)\s*\(", header)) - {"mcq_conv_params"}
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    lib = ctypes.CDLL(_lib.load().__dict__.get("_name", _lib.library_path()))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().mcq_version() == 2
    assert b"bad argument" in _lib.load().mcq_error_string(-1)
    # explicit options (no environment reads on the launch path): round trip, unknown names rejected
    assert _lib.get_option("pdl") == 1
    _lib.set_option("pdl", 0)
    assert _lib.get_option("pdl") == 0
    _lib.set_option("pdl", 1)
    with pytest.raises(RuntimeError):
        _lib.set_option("no_such_knob", 1)
    src = open(os.path.join(ROOT, "mcquic_b200", "csrc", "mcq_api.cu")).read()
    assert "getenv" not in src


def test_conv_params_struct_matches_header_field_order():
    header = open(os.path.join(ROOT, "include", "mcquic_b200.h")).read()
    body = header[header.index("typedef struct mcq_conv_params {"):header.index("} mcq_conv_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        first, *rest = decl.split(",")
        names.append(first.split()[-1].lstrip("*"))
        names += [r.strip().lstrip("*") for r in rest]
    assert names == [f[0] for f in _lib.ConvParams._fields_]


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    p = _lib.ConvParams()
    assert lib.mcq_conv2d(ctypes.byref(p), None) == -1                     # null operands
    assert lib.mcq_vq_assign(None, None, None, None, None, None, None, 1, 1, 1, 1, 1, 4, None) == -1
    assert lib.mcq_code_histogram(None, 1, 1, 1, 1, None, None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")


def test_state_dict_layout_is_the_references():
    sd = Compressor(32, 2, [16, 8]).state_dict()
    keys = set(sd)
    for must in ["_encoder.0.weight", "_encoder.2._branch.2.gamma", "_encoder.2._branch.2.beta_reparam.lowerBound.bound",
                 "_encoder.2._skip.weight", "_encoder.3._sideBranch.3.weight", "_decoder.1._branch.1.0.weight",
                 "_decoder.1._skip.0.bias", "_decoder.6.0.weight", "_quantizer._encoders.0._quantizer._codebook",
                 "_quantizer._encoders.0._dequantizer._codebook", "_quantizer._decoders.0._dequantizer._codebook",
                 "_quantizer._encoders.0._quantizer._temperature", "_quantizer._encoders.0._quantizer._bound.bound",
                 "_quantizer._entropyCoder._freqEMA.1", "_quantizer._encoders.0._latentHead.2.weight"]:
        assert must in keys, must
    assert "_quantizer._encoders.1._latentHead.2.weight" not in keys      # last level has no latentHead / sideHead
    assert "_quantizer._decoders.1._sideHead.1.weight" not in keys
    assert sd["_decoder.6.0.weight"].shape == (12, 32, 3, 3)
    assert sd["_quantizer._encoders.1._quantizer._codebook"].shape == (2, 8, 16)


def test_state_dict_matches_reference_exactly():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref = ref_import.build_reference_compressor(32, 2, [16, 8]).state_dict()
    mine = Compressor(32, 2, [16, 8]).state_dict()
    assert list(ref) == list(mine)
    assert all(ref[k].shape == mine[k].shape and ref[k].dtype == mine[k].dtype for k in ref)


def test_aligned_padding_amounts():
    assert aligned_pad_amounts(256, 256) == (0, 0, 256, 256)
    assert aligned_pad_amounts(200, 136) == (28, 60, 256, 256)
    assert aligned_pad_amounts(1152, 2048) == (0, 0, 1152, 2048)
    assert aligned_pad_amounts(129, 383) == (63, 0, 256, 384)
    for h, w in [(200, 136), (257, 300), (768, 512)]:
        top, left, hp, wp = aligned_pad_amounts(h, w)
        assert tuple(O.aligned_padding(torch.zeros(1, 3, h, w)).shape[-2:]) == (hp, wp)


def test_split_weight_is_accurate_and_pixel_shuffle_permutation():
    w = torch.randn(16, 40) * 0.03
    hi, lo, scale = split_weight(w)
    rec = (hi.double() + lo.double() / 2048.0) * scale
    assert float((rec - w.double()).abs().max()) <= float(w.abs().max()) * 2.0 ** -21
    wt = torch.randn(16, 8, 3, 3)
    pc = pack_conv(wt, torch.arange(16.0), 1, _lib.STORE_SHUFFLE_NHWC, "cpu")
    # GEMM column (2i+j)*cq + c must hold conv channel 4c + 2i + j
    assert pc.bias.tolist() == [float(4 * c + s) for s in range(4) for c in range(4)]
    pc2 = pack_conv(torch.randn(12, 64, 3, 3), torch.zeros(12), 1, _lib.STORE_SHUFFLE_NCHW, "cpu")
    assert pc2.cout == 12 and pc2.cout_pad == 16 and pc2.w_hi.shape == (16, 576)


@pytest.mark.parametrize("name,passes_err", [("compressor_small", 2e-6)])
def test_host_logic_against_oracle_through_emulated_abi(name, passes_err):
    g, cfg = load_golden(name)
    sd, x = golden_inputs(cfg)
    model = Compressor(cfg["channel"], cfg["m"], cfg["k"]).eval()
    model.load_state_dict(sd)
    model._engine = Engine(lib=EmulatedLib())
    ref = golden_codes(g, len(cfg["k"]))
    hist = torch.zeros(sum(cfg["m"] * k for k in cfg["k"]), dtype=torch.int32)
    codes = model.encode(x, hist=hist)
    flips, total, _ = code_report(codes, ref)
    assert flips == 0, f"{flips}/{total}"
    assert all(c.dtype == torch.int64 and c.is_contiguous() for c in codes)
    assert model.engine.lib.launches == 170 - 0  # one fused launch per conv/GDN + stem + VQ/dequant (no elementwise launches)
    # histogram fused into the VQ launch == oracle bincount
    exp = torch.cat([h.flatten() for h in O.code_histogram(ref, cfg["k"])]).int()
    assert torch.equal(hist, exp)
    xref = O.decode(sd, ref)
    model.decode_passes = 3
    assert float((model.decode(ref) - xref).abs().max()) <= passes_err
    model.decode_passes = 1
    assert float((model.decode(ref) - xref).abs().max()) <= 1e-3          # north_star pixel tolerance
    with pytest.raises(RuntimeError):
        model.decode([ref[0], ref[1]])                                     # wrong number of levels
    bad = [r.clone() for r in ref]
    bad[0][0, 0, 0, 0] = cfg["k"][0]
    with pytest.raises(RuntimeError):
        model.decode(bad)                                                  # code outside [0, k)
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(1, 1, 128, 128))                          # not an RGB batch


def test_quantizer_sub_api_through_emulated_abi():
    from mcquic_b200 import engine as E
    from mcquic_b200.modules.quantizer import _multiCodebookDeQuantization, _multiCodebookQuantization
    old = E._DEFAULT
    E._DEFAULT = Engine(lib=EmulatedLib())
    try:
        cb = torch.nn.Parameter(torch.randn(3, 50, 8) * 0.2)
        q, dq = _multiCodebookQuantization(cb), _multiCodebookDeQuantization(cb)
        x = torch.randn(2, 24, 5, 7) * 0.2
        code = q.encode(x)
        assert torch.equal(code, O.vq_assign(x, cb.data))
        assert torch.equal(dq.decode(code), O.vq_dequantize(code, cb.data))
        c2, logit = q.logits(x)
        assert torch.equal(c2, code)
        assert float((logit - O.vq_logits(x, cb.data, q._temperature.data)).abs().max()) < 1e-5
        sample, code2, onehot, logit2 = q(x)
        assert sample.shape == onehot.shape == logit2.shape == (2, 3, 5, 7, 50)
        assert torch.equal(onehot.argmax(-1), code2) and abs(float(sample.sum()) - 2 * 3 * 5 * 7) < 1e-3   # (1 - s) + s
        assert torch.allclose(dq(onehot), dq.decode(code2), atol=1e-6)
        assert float((sample - sample.round()).abs().max()) <= 2e-7 and float(sample.max()) <= 1.0 + 2e-7
    finally:
        E._DEFAULT = old


def test_product_refuses_cpu_tensors_and_missing_library(monkeypatch):
    model = Compressor(32, 1, [16, 8]).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model.encode(torch.zeros(1, 3, 128, 128))
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "library_path", lambda: "/nonexistent/libmcquic_b200.so")
    monkeypatch.setattr(_lib._build, "build", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no nvcc")))
    with pytest.raises(RuntimeError, match="missing|not found"):
        Engine()


def test_empty_batch_raises_like_the_reference():
    """upstream cannot encode an empty batch either: `_distance` reshapes 0 elements with a -1 (quantizer.py:158) ->
    RuntimeError; same exception type here, before anything is launched"""
    model = Compressor(32, 2, [16, 8, 4]).eval()
    model._engine = Engine(lib=EmulatedLib())
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(0, 3, 200, 136))
    assert model.engine.lib.launches == 0
    from mcquic_b200.utils.synthetic import synthetic_state_dict
    with pytest.raises(RuntimeError):
        O.encode(synthetic_state_dict(32, 2, [16, 8, 4], seed=0), torch.zeros(0, 3, 200, 136))


def test_pack_conv_pads_channels_the_kernels_cannot_address():
    """Neon's RGB-in / RGB-out / bias-free convs (compressor.py:188,227; quantizer.py:606): cin 3 -> 8 zero columns,
    plain-NHWC cout 3 -> 8 zero rows with zero bias, bias=None -> zero bias; values of the real part untouched"""
    w = torch.randn(16, 3, 3, 3)
    pc = pack_conv(w, torch.arange(16.0), 1, _lib.STORE_NHWC, "cpu")
    assert (pc.cin, pc.cout, pc.cout_pad) == (8, 16, 16) and pc.w_hi.shape == (16, 9 * 8)
    rec = ((pc.w_hi.double() + pc.w_lo.double() / 2048.0) * pc.w_scale).reshape(16, 3, 3, 8)
    assert float((rec[..., :3] - w.permute(0, 2, 3, 1).double()).abs().max()) <= float(w.abs().max()) * 2.0 ** -21
    assert float(rec[..., 3:].abs().max()) == 0.0
    pc = pack_conv(torch.randn(3, 64, 3, 3), torch.ones(3), 1, _lib.STORE_NHWC, "cpu")
    assert (pc.cin, pc.cout) == (64, 8) and pc.bias.tolist() == [1.0, 1.0, 1.0, 0, 0, 0, 0, 0]
    assert float(pc.w_hi[3:8].abs().max()) == 0.0
    pc = pack_conv(torch.randn(32, 8, 1, 1), None, 1, _lib.STORE_NHWC, "cpu")
    assert pc.bias.tolist() == [0.0] * 32 and pc.ksize == 1


def test_code_frequency_with_one_codebook_count_per_level():
    """VariousMCoder (entropyCoder.py:293-322): m is a list, EMA 0.998, flat histogram = level-major segments"""
    from mcquic_b200.modules.quantizer import CodeFrequency
    cf = CodeFrequency([1, 2], [4, 3], ema=0.998)
    assert [tuple(f.shape) for f in cf._freqEMA] == [(1, 4), (2, 3)] and cf.hist_size() == 4 + 6
    hist = torch.tensor([4, 0, 0, 0, 1, 1, 2, 0, 3, 0], dtype=torch.int32)
    cf.update(hist)
    want0 = 0.002 * torch.tensor([[1.0, 0, 0, 0]]) + 0.998 * 0.25
    want1 = 0.002 * torch.tensor([[.25, .25, .5], [0, 1.0, 0]]) + 0.998 / 3
    assert torch.allclose(cf._freqEMA[0], want0) and torch.allclose(cf._freqEMA[1], want1)
    assert CodeFrequency(2, [4, 3]).hist_size() == 14


def test_compressor_forward_values_through_emulated_abi():
    """BaseCompressor.forward (compressor.py:35-43), forward values only: shapes, finiteness, codes = argmax of the
    returned logits, level-0 codes = encode()'s (same deterministic logits), EMA updated"""
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    model = Compressor(32, 2, [16, 8]).eval()
    model.load_state_dict(synthetic_state_dict(32, 2, [16, 8], seed=0))
    model._engine = Engine(lib=EmulatedLib())
    x = uniform((2, 3, 64, 96), "forward.image", 2)
    before = model._quantizer._entropyCoder._freqEMA[0].clone()
    torch.manual_seed(1)
    xhat, yhat, codes, logits = model(x)
    assert tuple(xhat.shape) == (2, 3, 64, 96) and tuple(yhat.shape) == (2, 32, 8, 12)
    assert bool(torch.isfinite(xhat).all()) and [tuple(c.shape) for c in codes] == [(2, 2, 4, 6), (2, 2, 2, 3)]
    assert all(torch.equal(c, l.argmax(-1)) for c, l in zip(codes, logits))
    assert not torch.equal(before, model._quantizer._entropyCoder._freqEMA[0])
    xa = uniform((1, 3, 128, 128), "forward.image.aligned", 2)     # encode() pads to multiples of 128, forward() does not
    assert torch.equal(model(xa)[2][0], model.encode(xa)[0])
    with pytest.raises(RuntimeError):
        model(x[:, :, :60])


def test_uint8_images_take_the_reference_input_transform():
    """encode(uint8) == encode((u8 / 255 - 0.5) * 2) (demo.py:110-118) -- host logic on the CPU model of the C ABI; the
    last layer's uint8 store is the reference's DeTransform (vision.py:135-146)."""
    from mcquic_b200.engine import Act
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    from oracle import mcquic_oracle as O
    model = Compressor(32, 2, [16, 8]).eval()
    model.load_state_dict(synthetic_state_dict(32, 2, [16, 8], seed=0))
    model._engine = Engine(lib=EmulatedLib())
    u8 = ((uniform((2, 3, 100, 130), "u8.host", 0) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
    xf = (u8.float() / 255.0 - 0.5) * 2
    a, b = model.encode(u8), model.encode(xf)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    assert all(torch.equal(p, q) for p, q in zip(a, O.encode(model.state_dict(), xf)))
    eng = model.engine
    eng.passes = 1
    status = torch.zeros(1, dtype=torch.int32)
    y = model._quantizer.decode_act(eng, a, eng.needs_of(model._decoder[0]), status)
    y5 = eng.run_seq(list(model._decoder)[:6], y, eng.needs_of(model._decoder[6]))
    xf32 = eng.run(model._decoder[6], y5, set()).f32
    out8 = torch.empty(tuple(xf32.shape), dtype=torch.uint8)
    n, c, h, w = out8.shape
    eng.run(model._decoder[6], y5, set(), into=Act(n, h, w, c, f32=out8))
    assert torch.equal(out8, O.to_uint8(xf32))
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(1, 3, 64, 64, dtype=torch.int32))


def test_packing_cache_survives_reuse_of_a_module_id():
    """`Engine._packed` / `autograd._PACKS` are keyed by id(module).  Once a module is collected CPython hands its id -- and
    the allocator its weight's address, with version counter 0 again -- to the next layer created; the cache entry must then
    not be served (seen on the GPU box: a dgrad packing with cin = 128 returned for a 64-channel layer)."""
    import gc
    from torch import nn
    from mcquic_b200 import autograd as A
    eng = Engine(lib=EmulatedLib())

    def make(cin, cout):
        return nn.Conv2d(cin, cout, 3, padding=1)

    first = make(16, 32)
    ident = id(first)
    pc = eng._packed_for(first)
    pk = A._packs_for(first, first.weight, first.bias)
    pk.fwd = pc
    assert pc.cin == 16
    del first, pc
    gc.collect()
    reused = None
    keep = []
    for _ in range(2000):                    # a freed object's address comes back almost immediately
        cand = make(24, 8)
        if id(cand) == ident:
            reused = cand
            break
        keep.append(cand)
    if reused is None:
        pytest.skip("the interpreter did not reuse the module id")
    assert eng._packed_for(reused).cin == 24
    pk2 = A._packs_for(reused, reused.weight, reused.bias)
    assert pk2 is not pk and pk2.fwd is None


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle, on a bounded sample) runs without a GPU and
    pairs with our arm: same metric, unit, direction and config.workload; ranks other than 0 print nothing"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    env = dict(os.environ, RANK="0")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "MPix/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == bench.WORKLOAD
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    other = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, env=dict(os.environ, RANK="1"), timeout=600)
    assert other.returncode == 0 and other.stdout.strip() == ""
