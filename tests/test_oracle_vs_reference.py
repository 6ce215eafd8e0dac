"""Pins oracle/mcquic_oracle.py to the reference itself: the unmodified reference sources (imported from
/root/reference with third-party stubs, oracle/ref_import.py) and the restatement must agree BIT FOR BIT on CPU.
Skipped where the reference tree is absent (the GPU box) -- tests/test_golden.py covers that case."""
import pytest
import torch

from oracle import mcquic_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.mark.parametrize("channel,m,k,n,h,w", [
    (32, 2, [64, 32, 16], 2, 128, 128),
    (32, 1, [128, 64], 1, 100, 180),      # two levels, AlignedPadding active (-> 128 x 256)
    (128, 1, [8192, 2048, 512], 1, 256, 256),  # qp=1, BASELINE.json configs[0]
])
def test_encode_decode_bit_identical(channel, m, k, n, h, w):
    model = ref_import.build_reference_compressor(channel, m, k, seed=0)
    sd = model.state_dict()
    torch.manual_seed(0)
    x = torch.rand(n, 3, h, w) * 2 - 1
    with torch.inference_mode():
        ref_codes = model.encode(x)
        ref_x = model.decode(ref_codes)
    codes = O.encode(sd, x)
    assert len(codes) == len(ref_codes) == len(k)
    for a, b in zip(codes, ref_codes):
        assert a.dtype == b.dtype == torch.int64 and a.shape == b.shape
        assert torch.equal(a, b)
    assert torch.equal(O.decode(sd, codes), ref_x)


def test_quantizer_pieces_bit_identical():
    ref_import.load()
    from mcquic.modules.quantizer import _multiCodebookDeQuantization, _multiCodebookQuantization
    torch.manual_seed(1)
    cb = torch.nn.Parameter(torch.randn(3, 50, 8) * 0.2)
    q = _multiCodebookQuantization(cb, 0.0)
    dq = _multiCodebookDeQuantization(cb)
    x = torch.randn(2, 24, 5, 7) * 0.2
    with torch.inference_mode():
        assert torch.equal(q._distance(x), O.vq_distance(x, cb.data))
        code = q.encode(x)
        assert torch.equal(code, O.vq_assign(x, cb.data))
        assert torch.equal(dq.decode(code), O.vq_dequantize(code, cb.data))
        assert torch.equal(q._logit(x) * q._bound(q._temperature), O.vq_logits(x, cb.data, q._temperature.data))


def test_aligned_padding_matches_reference():
    ref_import.load()
    from mcquic.data.transforms import AlignedPadding
    pad = AlignedPadding()
    for h, w in [(256, 256), (200, 136), (129, 1), (1152, 2048), (127, 255), (128, 300)]:
        if h < 64 or w < 64:
            continue  # reflect padding needs pad < size
        x = torch.rand(1, 3, h, w)
        assert torch.equal(pad(x), O.aligned_padding(x))


def test_histogram_matches_bincount_and_onehot_sum():
    codes = [torch.randint(0, 16, (3, 2, 4, 4)), torch.randint(0, 8, (3, 2, 2, 2))]
    hist = O.code_histogram(codes, [16, 8])
    for code, k, h in zip(codes, [16, 8], hist):
        onehot = torch.zeros(*code.shape, k).scatter_(-1, code[..., None], 1)
        assert torch.equal(onehot.sum((0, 2, 3)).long(), h)   # entropyCoder.py:33
