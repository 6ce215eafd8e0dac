"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_import.py) on deterministic inputs.  Run here (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py

Inputs and weights are *not* stored: they come from the counter-hash generator in
mcquic_b200/utils/synthetic.py and are regenerated bit-identically by the tests.  Stored: the reference's
outputs (all code indices, pixels -- full for the small case, a strided sample + statistics for qp=1), the
top-2 distance margins of every code (so a flipped near-tie can be told from a bug) and sha256 digests.
TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402
from oracle import mcquic_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (channel, m, k, n, h, w, pixel sample stride)
CASES = {
    "compressor_small": (64, 2, [256, 128, 64], 2, 200, 136, 2),        # padded to 256x256 by AlignedPadding
    "compressor_qp1_256": (128, 1, [8192, 2048, 512], 1, 256, 256, 4),   # BASELINE.json configs[0]
    "compressor_c192_m6": (192, 6, [2048, 2048, 2048], 1, 128, 128, 4),  # "qp=3" shape of configs[2], one tile
}


def synthetic_image(n, h, w, seed):
    return uniform((n, 3, h, w), "image", seed)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_import.load()
    from mcquic.modules.compressor import Compressor
    from mcquic.modules.quantizer import _multiCodebookDeQuantization, _multiCodebookQuantization
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(8)
    for name, (c, m, k, n, h, w, stride) in CASES.items():
        sd = synthetic_state_dict(c, m, k, seed=0)
        model = Compressor(c, m, list(k)).eval()
        model.load_state_dict(sd)
        x = synthetic_image(n, h, w, seed=1)
        with torch.inference_mode():
            codes = model.encode(x)
            xhat = model.decode(codes)
        _, margins = O.encode(sd, x, with_margin=True)
        rec = {"config": np.array([c, m, n, h, w, stride] + list(k), dtype=np.int64),
               "xhat_sample": xhat[..., ::stride, ::stride].numpy().astype(np.float32),
               "xhat_stats": np.array([float(xhat.min()), float(xhat.max()), float(xhat.mean()), float(xhat.std())]),
               "xhat_sha256": np.array(sha(xhat)), "codes_sha256": np.array(sha(torch.cat([q.flatten() for q in codes])))}
        for lv, (q, mg) in enumerate(zip(codes, margins)):
            rec[f"codes_{lv}"] = q.numpy().astype(np.int32)
            rec[f"margin_{lv}"] = mg.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, [tuple(q.shape) for q in codes], tuple(xhat.shape), "min margin", min(float(mg.min()) for mg in margins))

    # ---- quantizer alone (BASELINE.json configs[2]: M=6, K=2048, d=32), reference module on seeded latents
    m, k, d, n, h, w = 6, 2048, 32, 4, 32, 32
    cb = uniform((m, k, d), "vq.codebook", 3) * ((2.0 / (5 * d)) ** 0.5 * 3 ** 0.5)
    x = uniform((n, m * d, h, w), "vq.latent", 3) * 0.26   # std 0.15 like observed latents
    q = _multiCodebookQuantization(torch.nn.Parameter(cb.clone()), 0.0)
    dq = _multiCodebookDeQuantization(q._codebook)
    with torch.inference_mode():
        code = q.encode(x)
        deq = dq.decode(code)
        logit = q._logit(x) * q._bound(q._temperature)
    np.savez_compressed(os.path.join(OUT, "vq_m6_k2048_d32.npz"),
                        config=np.array([m, k, d, n, h, w], dtype=np.int64), codes=code.numpy().astype(np.int32),
                        margin=O.vq_margin(x, cb).numpy().astype(np.float32),
                        deq_sha256=np.array(sha(deq)), logit_sample=logit[:, :, ::8, ::8, ::64].numpy())
    print("vq", tuple(code.shape), "min margin", float(O.vq_margin(x, cb).min()))


def blocks():
    """tests/golden/blocks_dense.npz: the reference's own ResidualBlock / AttentionBlock classes (GroupNorm via
    denseNorm=True, conv1x1 skips) on the deterministic weights / inputs of tests/common.py:DENSE_BLOCKS."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import DENSE_BLOCKS, dense_block_inputs, dense_stride
    ref_import.load()
    import mcquic.nn.blocks as RB
    torch.set_num_threads(8)
    rec = {}
    for name, (kind, args, shape) in DENSE_BLOCKS.items():
        block, x = dense_block_inputs(name, getattr(RB, kind))
        with torch.inference_mode():
            y = block(x)
        rec[name] = y[..., ::dense_stride(shape), ::dense_stride(shape)].numpy().astype(np.float32)
        rec[name + "_sha256"] = np.array(sha(y))
        print(name, tuple(y.shape), "absmax", float(y.abs().max()))
    np.savez_compressed(os.path.join(OUT, "blocks_dense.npz"), **rec)


def neon():
    """tests/golden/neon_*.npz: the reference's own `Neon` (compressor.py:181-233) on tests/common.py:NEON_CASES."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import NEON_CASES, neon_inputs
    ref_import.load()
    from mcquic.modules.compressor import Neon
    torch.set_num_threads(8)
    for name, (c, k, size, dense, n, h, w) in NEON_CASES.items():
        model, x = neon_inputs(name, Neon)
        with torch.inference_mode():
            codes = model.encode(x)
            xhat = model.decode(codes)
        _, margins = O.neon_encode(model.state_dict(), x, size, with_margin=True)
        rec = {"xhat": xhat.numpy().astype(np.float32), "xhat_sha256": np.array(sha(xhat)),
               "codes_sha256": np.array(sha(torch.cat([q.flatten() for q in codes])))}
        for j, (q, mg) in enumerate(zip(codes, margins)):
            rec[f"codes_{j}"] = q.numpy().astype(np.int32)
            rec[f"margin_{j}"] = mg.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, [tuple(q.shape) for q in codes], tuple(xhat.shape), "min margin", min(float(mg.min()) for mg in margins),
              "xhat absmax", float(xhat.abs().max()))


def _codes_rec(codes, margins):
    """codes as int16 (k <= 8192 everywhere), the oracle's top-2 relative margins next to them"""
    rec = {"codes_sha256": np.array(sha(torch.cat([q.flatten() for q in codes])))}
    for lv, (q, mg) in enumerate(zip(codes, margins)):
        assert int(q.max()) < 32768
        rec[f"codes_{lv}"] = q.numpy().astype(np.int16)
        rec[f"margin_{lv}"] = mg.numpy().astype(np.float32)
    rec["min_margin"] = np.array(min(float(mg.min()) for mg in margins))
    return rec


def baseline_configs():
    """The BASELINE.json configurations themselves, from the imported reference:
    * tests/golden/bench_qp1_n64.npz: ALL codes of the 64 images `bench.py` times (configs[1]: qp=1, 64x3x256x256,
      input = uniform(.., "bench.image.0", 0)), pixels sampled every 16th row / column + sha256;
    * tests/golden/compressor_q6_512.npz: Q6 = Compressor(192, 6, [2048]*3) on 10 images of 512x512 (configs[2]'s model);
    * tests/golden/vq_cfg3_full.npz: configs[2]'s VQ alone at FULL size (N=32, 32x32 grid, M=6, K=2048, d=32), the
      reference `_multiCodebookQuantization.encode` run image by image (images are independent in the bmm)."""
    ref_import.load()
    from mcquic.modules.compressor import Compressor
    from mcquic.modules.quantizer import _multiCodebookQuantization
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(8)

    def run(c, m, k, x, chunk, stride, with_pixels=True):
        sd = synthetic_state_dict(c, m, k, seed=0)
        model = Compressor(c, m, list(k)).eval()
        model.load_state_dict(sd)
        codes, margins, xs = [], [], []
        for i in range(0, x.shape[0], chunk):
            with torch.inference_mode():
                cd = model.encode(x[i:i + chunk])
                if with_pixels:
                    xs.append(model.decode(cd)[..., ::stride, ::stride].clone())
            # the oracle must agree with the reference bit for bit here as well (it supplies the margins)
            oc, mg = O.encode(sd, x[i:i + chunk], with_margin=True)
            assert all(torch.equal(a, b) for a, b in zip(cd, oc)), "oracle != reference"
            codes.append(cd)
            margins.append(mg)
            print("  images", i, "..", i + chunk, flush=True)
        codes = [torch.cat([cc[lv] for cc in codes]) for lv in range(len(k))]
        margins = [torch.cat([mm[lv] for mm in margins]) for lv in range(len(k))]
        rec = _codes_rec(codes, margins)
        if with_pixels:
            rec["xhat_sample"] = torch.cat(xs).numpy().astype(np.float32)
        rec["config"] = np.array([c, m, x.shape[0], x.shape[2], x.shape[3], stride] + list(k), dtype=np.int64)
        return rec

    x = uniform((64, 3, 256, 256), "bench.image.0", 0)
    rec = run(128, 1, [8192, 2048, 512], x, 8, 16)
    np.savez_compressed(os.path.join(OUT, "bench_qp1_n64.npz"), **rec)
    print("bench_qp1_n64", sum(rec[f"codes_{lv}"].size for lv in range(3)), "codes, min margin", float(rec["min_margin"]))

    x = uniform((10, 3, 512, 512), "q6.image", 5)
    rec = run(192, 6, [2048, 2048, 2048], x, 1, 32)
    np.savez_compressed(os.path.join(OUT, "compressor_q6_512.npz"), **rec)
    print("compressor_q6_512", sum(rec[f"codes_{lv}"].size for lv in range(3)), "codes, min margin", float(rec["min_margin"]))

    m, k, d, n, h, w = 6, 2048, 32, 32, 32, 32
    cb = uniform((m, k, d), "vq.codebook", 3) * ((2.0 / (5 * d)) ** 0.5 * 3 ** 0.5)
    x = uniform((n, m * d, h, w), "vq.latent.full", 3) * 0.26
    q = _multiCodebookQuantization(torch.nn.Parameter(cb.clone()), 0.0)
    with torch.inference_mode():
        code = torch.cat([q.encode(x[i:i + 1]) for i in range(n)])
    margin = torch.cat([O.vq_margin(x[i:i + 1], cb) for i in range(n)])
    assert torch.equal(code, torch.cat([O.vq_assign(x[i:i + 1], cb) for i in range(n)]))
    np.savez_compressed(os.path.join(OUT, "vq_cfg3_full.npz"), config=np.array([m, k, d, n, h, w], dtype=np.int64),
                        codes=code.numpy().astype(np.int16), margin=margin.numpy().astype(np.float16),
                        min_margin=np.array(float(margin.min())), codes_sha256=np.array(sha(code)))
    print("vq_cfg3_full", code.numel(), "codes, min margin", float(margin.min()))


def train():
    """tests/golden/train_*.npz: ONE training step of the reference's own `Neon` (BaseCompressor.forward in training mode,
    compressor.py:35-43 -> ResidualBackwardQuantizer.forward, quantizer.py:727-765) with an MSE loss, on CPU in fp32:
    loss, xHat, codes, the updated frequency EMA and, for every parameter, the gradient's L2 norm plus a strided sample.
    The Gumbel / drop uniforms come from tests/common.py:DeterministicRand so the CUDA path can draw the same ones."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import TRAIN_CASES, DeterministicRand, grad_sample, train_inputs
    import torch.distributed as dist
    ref_import.load()
    from mcquic.modules.compressor import Neon
    torch.set_num_threads(8)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if not dist.is_initialized():      # EntropyCoder.forward all-reduces unconditionally (entropyCoder.py:314)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    for name in TRAIN_CASES:
        model, x = train_inputs(name, Neon)
        model.train()
        with DeterministicRand(name) as rnd:
            xHat, yHat, codes, logits = model(x.clone())
            loss = torch.nn.functional.mse_loss(xHat, x)
            loss.backward()
            calls = rnd.calls
        rec = {"loss": np.array(float(loss)), "xhat": xHat.detach().numpy().astype(np.float32),
               "yhat": yHat.detach().numpy().astype(np.float32), "rand_calls": np.array(calls)}
        for j, (c, lg) in enumerate(zip(codes, logits)):
            rec[f"codes_{j}"] = c.numpy().astype(np.int32)
            top2 = torch.topk(lg.detach(), 2, dim=-1).values
            rec[f"logit_margin_{j}"] = np.array(float((top2[..., 0] - top2[..., 1]).min()))
        for j, f in enumerate(model._quantizer._entropyCoder._freqEMA):
            rec[f"freq_{j}"] = f.detach().numpy().astype(np.float32)
        seen = set()
        missing = []
        for key, p in model.named_parameters():
            if p.data_ptr() in seen:
                continue
            seen.add(p.data_ptr())
            if p.grad is None:
                missing.append(key)
                continue
            rec["gnorm." + key] = np.array(float(p.grad.norm()))
            rec["gsample." + key] = grad_sample(p.grad).numpy().astype(np.float32)
        rec["no_grad_params"] = np.array(";".join(missing))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "loss", float(loss), "params with grad", sum(1 for k_ in rec if k_.startswith("gnorm.")),
              "without", len(missing), "rand calls", calls, [tuple(c.shape) for c in codes])


CONTAINER_CASES = [
    # (version, qp, m, heights, widths, k, (height, width, channel), contents)
    ("0.1.40", "qp_1_msssim", [1, 1, 1], [16, 8, 4], [16, 8, 4], [8192, 2048, 512], (256, 256, 3),
     [b"\x01\x02", b"\x03", b"\x04\x05\x06"]),
    ("0.1.40", "qp_2_msssim", [2, 2, 2], [72, 36, 18], [120, 60, 30], [8192, 2048, 512], (1100, 1900, 3),
     [bytes(range(256)) * 40, bytes(range(200)), b"\xff" * 33]),
    ("0.1.3", "-1", [6, 6, 6], [32, 16, 8], [32, 16, 8], [2048, 2048, 2048], (512, 512, 3),
     [b"a" * 70000, b"\x00" * 300, b"\xc4"]),                      # bin8 / bin16 / bin32 length prefixes
]


def container():
    """tests/golden/container_reference.json: the bytes the reference's OWN `File.serialize` (specification.py:147-149,
    run unmodified through the functional marshmallow stand-in of oracle/ref_import.py) produces for CONTAINER_CASES."""
    import json
    ref_import.load()
    from mcquic.utils.specification import CodeSize, File, FileHeader, ImageSize
    rec = []
    for version, qp, m, hs, ws, k, (ih, iw, ic), contents in CONTAINER_CASES:
        f = File(FileHeader(version, qp, CodeSize(m, hs, ws, k), ImageSize(ih, iw, ic)), list(contents))
        data = f.serialize()
        back = File.deserialize(data)
        assert back.fileHeader.codeSize.k == k and list(back.contents) == list(contents)
        rec.append({"sha256": hashlib.sha256(data).hexdigest(), "size": len(data), "head_hex": data[:160].hex(),
                    "bpp": f.BPP, "str": str(f)})
        print("container", qp, len(data), "bytes")
    with open(os.path.join(OUT, "container_reference.json"), "w") as fp:
        json.dump(rec, fp, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--container" in sys.argv:
        container()
        sys.exit(0)
    if "--train" in sys.argv:
        train()
        sys.exit(0)
    if "--blocks" in sys.argv:
        blocks()
    elif "--neon" in sys.argv:
        neon()
    elif "--baseline" in sys.argv:
        baseline_configs()
    else:
        main()
        blocks()
        neon()
        baseline_configs()
        container()
        train()
