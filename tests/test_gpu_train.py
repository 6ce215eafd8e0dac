"""Training-step kernels (SURVEY.md section 8f NEXT-3) through the C ABI: forward, input gradient (dgrad) and weight
gradient (wgrad, csrc/conv_wgrad.cuh) of the convolutions against torch autograd of F.conv2d evaluated in fp64.
One fp16 pass with fp32 accumulation = TF32-grade (what the reference trains with): tolerance 3e-3 of the largest value."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from mcquic_b200 import autograd as A
from mcquic_b200.utils.synthetic import uniform

pytestmark = pytest.mark.gpu

TOL = 3e-3


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max()) / (float(b.abs().max()) + 1e-300)


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,bias,gscale", [
    (2, 16, 16, 128, 128, 3, 1, True, 1.0),       # the common block convolution
    (2, 32, 32, 64, 128, 3, 2, True, 1e-7),        # strided (ResidualBlockWithStride), tiny gradients: device-side scaling
    (3, 8, 8, 128, 128, 1, 1, True, 1.0),          # 1x1 (gate / GDN)
    (2, 16, 16, 32, 32, 3, 1, True, 1e3),          # Neon's quantizer nets: partly filled 64-channel chunks
    (1, 24, 40, 8, 32, 1, 1, False, 1.0),          # bias-free 1x1, 8 channels, ragged pixel tiles
    (2, 32, 32, 3, 64, 3, 1, True, 1.0),           # RGB stem of Neon (input padded to 8 channels; no input gradient needed)
    (2, 16, 16, 64, 3, 3, 1, True, 1e-5),          # Neon's last layer C -> 3
    (2, 32, 32, 128, 512, 3, 1, True, 1.0),        # the convolution in front of PixelShuffle
    (2, 12, 20, 256, 256, 3, 1, True, 1.0),        # 256 channels (a800_16 trunk): 2 x 2 channel tiles
    (1, 64, 64, 128, 128, 3, 2, True, 1.0),
    (5, 4, 4, 128, 128, 3, 1, True, 1.0),          # 4x4 maps: a pixel tile spans 8 images, ragged in n
])
def test_conv_function_matches_fp64_autograd(n, h, w, cin, cout, k, stride, bias, gscale):
    conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(uniform(tuple(conv.weight.shape), "train.w", 3) / (cin * k * k) ** 0.5)
        if bias:
            conv.bias.copy_(uniform((cout,), "train.b", 3) * 0.1)
    x = uniform((n, cin, h, w), "train.x", 4)
    g = uniform((n, cout, h // stride, w // stride), "train.g", 5) * gscale
    # fp64 reference
    xr = x.double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    br = conv.bias.detach().double().requires_grad_(True) if bias else None
    yr = F.conv2d(xr, wr, br, stride=stride, padding=k // 2)
    yr.backward(g.double())
    conv = conv.cuda()
    assert A.conv_supported(conv)
    xg = x.cuda().requires_grad_(True)
    before = A.train_engine().lib.mcq_kernel_launch_count()
    y = A.conv2d(conv, xg)
    assert tuple(y.shape) == tuple(yr.shape)
    y.backward(g.cuda())
    torch.cuda.synchronize()
    assert A.train_engine().lib.mcq_kernel_launch_count() - before >= 6      # split + conv, split + dgrad, wgrad + reduce
    assert _rel(y, yr.detach()) <= TOL
    assert _rel(xg.grad, xr.grad) <= TOL
    assert _rel(conv.weight.grad, wr.grad) <= TOL
    if bias:
        assert _rel(conv.bias.grad, br.grad) <= 1e-5
    assert A.train_engine().lib.mcq_device_error_flag() == 0


def test_unsupported_convolutions_use_the_library_path_and_still_differentiate():
    conv = nn.Conv2d(32, 32, 3, stride=2, padding=1).cuda()       # stride 2 with a 32-channel tap view
    assert not A.conv_supported(conv)
    x = uniform((2, 32, 16, 16), "train.x32", 1).cuda().requires_grad_(True)
    A.conv2d(conv, x).sum().backward()
    assert x.grad is not None and conv.weight.grad is not None


# ---------------------------------------------------------------------------------------------------------------------
# one whole training step: BaseCompressor.forward in training mode + MSE loss + backward
import numpy as np  # noqa: E402

from common import GOLDEN, TRAIN_CASES, DeterministicRand, grad_sample, log_parity, train_inputs  # noqa: E402


def _step(model, x, tag):
    model.zero_grad(set_to_none=True)
    with DeterministicRand(tag) as rnd:
        xHat, yHat, codes, logits = model(x)
        loss = F.mse_loss(xHat, x)
        loss.backward()
    torch.cuda.synchronize()
    return xHat, yHat, codes, logits, loss, rnd.calls


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_training_step_matches_the_reference_golden(name):
    """the reference's own Neon, one forward + backward on CPU in fp32 (tests/golden/train_*.npz, oracle/gen_golden.py
    --train) vs the CUDA path with fp32-grade forward / dgrad convolutions (3 passes) under the same uniforms: identical
    codes, loss to 1e-4, xHat to 1e-4 of its range, frequency EMA, and every parameter's gradient (norm to 1 %, a strided
    sample to 1 % of its largest entry; the weight gradients are one fp16 pass = TF32-grade)."""
    from mcquic_b200 import Neon
    g = np.load(f"{GOLDEN}/{name}.npz")
    model, x = train_inputs(name, Neon)
    model = model.cuda().train()
    A.set_passes(3)
    try:
        before = A.train_engine().lib.mcq_kernel_launch_count()
        xHat, yHat, codes, logits, loss, calls = _step(model, x.cuda(), name)
        launches = A.train_engine().lib.mcq_kernel_launch_count() - before
    finally:
        A.set_passes(1)
    assert calls == int(g["rand_calls"]) and launches > 300
    flips = total = 0
    for j, c in enumerate(codes):
        ref = torch.from_numpy(g[f"codes_{j}"].astype(np.int64))
        flips += int((c.cpu() != ref).sum())
        total += ref.numel()
    log_parity(f"golden {name} (training forward, sampled codes)", flips, total)
    assert flips == 0
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    xr = torch.from_numpy(g["xhat"])
    assert float((xHat.detach().cpu() - xr).abs().max()) <= 1e-4 * float(xr.abs().max())
    yr = torch.from_numpy(g["yhat"])
    assert float((yHat.detach().cpu() - yr).abs().max()) <= 1e-4 * float(yr.abs().max())
    for j, f in enumerate(model._quantizer._entropyCoder._freqEMA):
        assert torch.allclose(f.detach().cpu(), torch.from_numpy(g[f"freq_{j}"]), rtol=0, atol=1e-7)
    # gradients that are zero in exact arithmetic (a convolution bias in front of a GroupNorm with one channel per group)
    # are rounding noise ~1e-9 on both sides: the absolute floor is 1e-6 of the largest gradient norm of the model
    floor = 1e-6 * max(float(g[k_]) for k_ in g.files if k_.startswith("gnorm."))
    seen, checked, worst = set(), 0, 0.0
    for key, p in model.named_parameters():
        if p.data_ptr() in seen:
            continue
        seen.add(p.data_ptr())
        if "gnorm." + key not in g.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, key
            continue
        assert p.grad is not None, key
        rn = float(g["gnorm." + key])
        assert abs(float(p.grad.norm()) - rn) <= 1e-2 * rn + floor, (key, float(p.grad.norm()), rn)
        rs = torch.from_numpy(g["gsample." + key])
        err = float((grad_sample(p.grad).cpu() - rs).abs().max())
        if float(rs.abs().max()) > floor:
            worst = max(worst, err / float(rs.abs().max()))
        assert err <= 1e-2 * float(rs.abs().max()) + floor, (key, err, float(rs.abs().max()))
        checked += 1
    assert checked > 400
    print(f"[train parity] {name}: {checked} parameter gradients checked, worst sample error {worst:.2e} of its largest entry")
    assert A.train_engine().lib.mcq_device_error_flag() == 0


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_training_step_one_pass(name):
    """the default (one fp16 pass everywhere = TF32-grade, like the reference's own training numerics): the loss stays within
    1e-2 of the fp32 reference's and the gradients point the same way (cosine > 0.98 over all sampled entries); hard
    sampling decisions may differ at near-ties, so no bit-level claim here."""
    from mcquic_b200 import Neon
    g = np.load(f"{GOLDEN}/{name}.npz")
    model, x = train_inputs(name, Neon)
    model = model.cuda().train()
    xHat, yHat, codes, logits, loss, calls = _step(model, x.cuda(), name)
    assert abs(float(loss) - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    a, b, seen = [], [], set()
    for key, p in model.named_parameters():
        if p.data_ptr() in seen or "gsample." + key not in g.files:
            continue
        seen.add(p.data_ptr())
        assert p.grad is not None and torch.isfinite(p.grad).all(), key
        a.append(grad_sample(p.grad).cpu().double())
        b.append(torch.from_numpy(g["gsample." + key]).double())
    a, b = torch.cat(a), torch.cat(b)
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    print(f"[train parity] {name} one pass: loss {float(loss):.6f} vs {float(g['loss']):.6f}, gradient cosine {cos:.5f}")
    assert cos > 0.98


def test_compressor_training_step_equals_a_library_evaluation_of_the_same_graph(monkeypatch):
    """`Compressor` cannot train upstream at HEAD (UMGMQuantizer hands a float where the frequency tensor is expected,
    quantizer.py:399 / SURVEY 8a row a14), so there is no reference golden for it: here the CUDA path (tcgen05 convolutions,
    VQ-launch logits) is compared with the SAME differentiable graph evaluated through torch's convolution and a plain torch
    distance, under the same uniforms."""
    from mcquic_b200 import Compressor
    from mcquic_b200.utils.synthetic import synthetic_state_dict
    sd = synthetic_state_dict(64, 2, [64, 32], seed=0)
    x = uniform((2, 3, 64, 64), "train.compressor.image", 1).cuda()

    def run(patched):
        model = Compressor(64, 2, [64, 32])
        model.load_state_dict(sd)
        model = model.cuda().train()
        if patched:
            monkeypatch.setattr(A, "conv2d", lambda conv, t: F.conv2d(t, conv.weight, conv.bias, conv.stride, conv.padding))
            monkeypatch.setattr(A, "conv2d_weights", lambda t, w, b, stride=1, owner=None: F.conv2d(t, w, b, stride, w.shape[-1] // 2))

            class Logits:
                @staticmethod
                def apply(t, cb, temp):
                    n, c, h, w = t.shape
                    m, k, d = cb.shape
                    xs = t.reshape(n, m, d, h * w).permute(0, 1, 3, 2)
                    dist = (xs ** 2).sum(-1, keepdim=True) + (cb ** 2).sum(-1)[None, :, None, :] - 2 * xs @ cb.transpose(1, 2)[None]
                    return (-dist / k ** 0.5 * temp.reshape(1, m, 1, 1)).reshape(n, m, h, w, k)
            monkeypatch.setattr(A, "_LogitsFn", Logits)
        out = _step(model, x, "train.compressor")
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        monkeypatch.undo()
        return out, grads

    A.set_passes(3)
    try:
        (xh, yh, codes, logits, loss, _), grads = run(False)
    finally:
        A.set_passes(1)
    (xr, yr, cr, lr, lossr, _), gref = run(True)
    assert all(torch.equal(a, b) for a, b in zip(codes, cr))
    assert abs(float(loss) - float(lossr)) <= 1e-4 * abs(float(lossr))
    assert set(grads) == set(gref) and len(grads) > 300
    for key in gref:
        scale = float(gref[key].abs().max()) + 1e-30
        assert float((grads[key] - gref[key]).abs().max()) <= 1e-2 * scale + 1e-12, key
