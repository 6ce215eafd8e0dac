"""Layer-chain launches (mcq_conv_chain, csrc/conv_chain.cuh) against the same layers issued one by one
(mcq_conv2d): the header promises bit-identical results, for any batch size (clusters own image groups; ragged
last group), both precisions, with and without CUDA graphs."""
import pytest
import torch

from mcquic_b200 import Compressor, _lib
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform

pytestmark = pytest.mark.gpu


def _model(channel, m, k):
    model = Compressor(channel, m, k).eval()
    model.load_state_dict(synthetic_state_dict(channel, m, k, seed=0))
    return model.cuda()


@pytest.fixture
def per_tap_kernels_only():
    """the chain kernel shares its tile body (and so its summation order) with the per-tap kernel conv_tc.cuh; the halo /
    CTA-pair kernels sum K in another order, so bit-exactness is checked with those switched off"""
    old = {name: _lib.get_option(name) for name in ("halo", "pair")}
    _lib.set_option("halo", 0)
    _lib.set_option("pair", 0)
    yield
    for name, v in old.items():
        _lib.set_option(name, v)


@pytest.mark.parametrize("n", [1, 3, 17, 64])
def test_chain_equals_layer_by_layer(n, per_tap_kernels_only):
    k = [8192, 2048, 512]
    model = _model(128, 1, k)
    model.use_graphs = False
    x = uniform((n, 3, 256, 256), "chain.image", 2).cuda()
    eng = model.engine
    out = {}
    for chain in (False, True):
        eng.chain = chain
        before = _lib.launch_count()
        codes = model.encode(x)
        launches_enc = _lib.launch_count() - before
        res = [codes]
        for passes in (1, 3):
            model.decode_passes = passes
            res.append(model.decode(codes))
        out[chain] = (res, launches_enc)
        assert eng.lib.mcq_device_error_flag() == 0
    (c0, a0, b0), l0 = out[False]
    (c1, a1, b1), l1 = out[True]
    assert all(torch.equal(u, v) for u, v in zip(c0, c1))
    assert torch.equal(a0, a1) and torch.equal(b0, b1)
    assert l1 < l0 // 2, (l0, l1)           # the tail really went through chains


def test_chain_under_graph_replay_and_other_shape(per_tap_kernels_only):
    model = _model(64, 2, [256, 128, 64])
    x = uniform((5, 3, 128, 384), "chain.image2", 3).cuda()
    model.engine.chain = False
    model.use_graphs = False
    ref = model.encode(x)
    xr = model.decode(ref)
    model.engine.chain = True
    model.use_graphs = True
    for _ in range(2):
        codes = model.encode(x)
        xh = model.decode(codes)
    assert all(torch.equal(u, v) for u, v in zip(ref, codes)) and torch.equal(xr, xh)
