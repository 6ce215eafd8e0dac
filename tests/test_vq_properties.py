"""Property tests (hypothesis) of the VQ oracle and of the product's host-side VQ plumbing, on CPU.

The reference carries only an author's comment that its bmm form of the distance equals the naive one
("ALREADY CHECKED CONSISTENCY WITH NAIVE IMPL.", mcquic/modules/quantizer.py:152,261) and no test.
Here: the oracle's restatement against a naive per-codebook evaluation in float64 over random shapes (k not a multiple of
any tile, d in {4..128}), exact ties -> first index like torch.argmin, de-quantisation = gather, logits = -distance /
sqrt(k) * max(temperature, eps); and the product's `_multiCodebookQuantization` / `_multiCodebookDeQuantization`
through the CPU model of the C ABI against the oracle on the same random cases.
"""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from emulator import EmulatedLib
from mcquic_b200 import engine as E
from mcquic_b200.engine import Engine
from mcquic_b200.modules.quantizer import _multiCodebookDeQuantization, _multiCodebookQuantization
from oracle import mcquic_oracle as O

shapes = st.tuples(st.integers(1, 3), st.integers(1, 4), st.integers(1, 5), st.integers(1, 5),
                   st.integers(1, 70), st.sampled_from([4, 8, 16, 32, 64, 128]))


def _case(n, m, h, w, k, d, seed):
    g = torch.Generator().manual_seed(seed)
    cb = torch.randn(m, k, d, generator=g) * (2.0 / (5 * d)) ** 0.5
    x = torch.randn(n, m * d, h, w, generator=g) * 0.15
    return x, cb


def _naive_distance(x, cb):
    """[n, m, h, w, k] squared L2 distances, float64, straight from the definition"""
    n, _, h, w = x.shape
    m, k, d = cb.shape
    xs = x.double().reshape(n, m, d, h, w).permute(0, 1, 3, 4, 2)            # [n, m, h, w, d]
    return ((xs[..., None, :] - cb.double()[None, :, None, None]) ** 2).sum(-1)


@settings(max_examples=40, deadline=None)
@given(shapes, st.integers(0, 2 ** 16))
def test_oracle_distance_equals_the_naive_definition(shape, seed):
    n, m, h, w, k, d = shape
    x, cb = _case(n, m, h, w, k, d, seed)
    dist = O.vq_distance(x, cb)
    naive = _naive_distance(x, cb)
    assert tuple(dist.shape) == (n, m, h, w, k)
    assert float((dist.double() - naive).abs().max()) <= 1e-5 * max(1.0, float(naive.max()))
    code = O.vq_assign(x, cb)
    # wherever the naive top-2 gap is resolvable in fp32, the codes are the naive argmin
    top2 = torch.topk(naive, min(2, k), dim=-1, largest=False).values
    clear = torch.ones_like(code, dtype=torch.bool) if k == 1 else (top2[..., 1] - top2[..., 0]) > 1e-5 * top2[..., 1].abs()
    assert torch.equal(code[clear], naive.argmin(-1)[clear])
    deq = O.vq_dequantize(code, cb)
    pick = cb[torch.arange(m)[None, :, None, None].expand_as(code), code]    # [n, m, h, w, d]
    assert torch.equal(deq, pick.permute(0, 1, 4, 2, 3).reshape(n, m * d, h, w))


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(0, 2 ** 16))
def test_exact_ties_resolve_to_the_first_index(shape, seed):
    n, m, h, w, k, d = shape
    if k < 2:
        k = 2
    x, cb = _case(n, m, h, w, k, d, seed)
    cb[:, k // 2] = cb[:, 0]                       # duplicate codeword: bit-identical distances at indices 0 and k//2
    code = O.vq_assign(x, cb)
    assert not bool((code == k // 2).any())
    old, E._DEFAULT = E._DEFAULT, Engine(lib=EmulatedLib())
    try:
        got = _multiCodebookQuantization(torch.nn.Parameter(cb)).encode(x)
    finally:
        E._DEFAULT = old
    assert not bool((got == k // 2).any())


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(0, 2 ** 16), st.floats(1e-8, 4.0))
def test_product_host_path_against_oracle(shape, seed, temperature):
    n, m, h, w, k, d = shape
    x, cb = _case(n, m, h, w, k, d, seed)
    old, E._DEFAULT = E._DEFAULT, Engine(lib=EmulatedLib())
    try:
        q = _multiCodebookQuantization(torch.nn.Parameter(cb.clone()))
        dq = _multiCodebookDeQuantization(q._codebook)
        with torch.no_grad():
            q._temperature.fill_(temperature)
        code, logit = q.logits(x)
        deq = dq.decode(code)
    finally:
        E._DEFAULT = old
    margin = O.vq_margin(x, cb) if k > 1 else torch.ones(n, m, h, w)
    ref = O.vq_assign(x, cb)
    assert code.dtype == torch.int64 and tuple(code.shape) == (n, m, h, w)
    assert torch.equal(code[margin > 1e-5], ref[margin > 1e-5])
    want = O.vq_logits(x, cb, torch.full((m, 1, 1, 1), temperature))
    assert float((logit - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    assert torch.equal(deq, O.vq_dequantize(code, cb))
    # logits = -distance / sqrt(k) * max(temperature, 1e-6)  (quantizer.py:181-183,204; base.py:31-54)
    scale = max(temperature, 1e-6) / k ** 0.5
    assert float((want + O.vq_distance(x, cb) * scale).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))


def test_histogram_is_the_one_hot_sum_for_random_codes():
    g = np.random.default_rng(0)
    for _ in range(10):
        n, m, h, w, k = (int(v) for v in g.integers(1, 6, 5))
        k += 1
        code = torch.from_numpy(g.integers(0, k, (n, m, h, w)))
        onehot = torch.zeros(n, m, h, w, k).scatter_(-1, code[..., None], 1)
        assert torch.equal(O.code_histogram([code], [k])[0], onehot.sum((0, 2, 3)).long())
