"""Shared test utilities (golden loading, deterministic inputs)."""
import os

import numpy as np
import torch

from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = g["config"].tolist()
    c, m, n, h, w, stride = cfg[:6]
    k = cfg[6:]
    return g, dict(channel=c, m=m, k=k, n=n, h=h, w=w, stride=stride)


def golden_inputs(cfg):
    sd = synthetic_state_dict(cfg["channel"], cfg["m"], cfg["k"], seed=0)
    x = uniform((cfg["n"], 3, cfg["h"], cfg["w"]), "image", 1)
    return sd, x


def golden_codes(g, levels):
    return [torch.from_numpy(g[f"codes_{lv}"].astype(np.int64)) for lv in range(levels)]


# every oracle / golden comparison of code indices reports here; tests/conftest.py prints the list at the end of the run,
# so the driver's log shows how many indices flipped (the green run must show 0 everywhere) and how thin the oracle's
# smallest top-2 margin was for that case
PARITY_LOG = []


def log_parity(case, flips, total, at=(), min_margin=None):
    PARITY_LOG.append(dict(case=case, flips=int(flips), total=int(total), flip_margins=[float(a) for a in at][:8],
                           min_margin=None if min_margin is None else float(min_margin)))
    print(f"[parity] {case}: flips {flips}/{total}" + ("" if min_margin is None else f", oracle min top-2 margin {min_margin:.3e}")
          + (f", margins at the flips {sorted(float(a) for a in at)[:8]}" if flips else ""))


def code_report(got, ref, margins=None):
    """(#mismatches, total, margins at the mismatching positions)"""
    flips, total, at = 0, 0, []
    for lv, (a, b) in enumerate(zip(got, ref)):
        mism = a.cpu() != b
        flips += int(mism.sum())
        total += b.numel()
        if margins is not None:
            at += torch.as_tensor(margins[lv])[mism].tolist()
    return flips, total, at


# ResidualBlock / AttentionBlock variants only the `Neon` tokenizer builds (GroupNorm via denseNorm=True, channel-
# changing blocks with a conv1x1 skip): name -> (kind, constructor args, input shape [n, c, h, w])
DENSE_BLOCKS = {
    "rb64_gn8": ("ResidualBlock", (64, 64, 8, True), (2, 64, 24, 20)),
    "rb64to128_gn32": ("ResidualBlock", (64, 128, 32, True), (2, 64, 16, 16)),
    "rb128to64_plain": ("ResidualBlock", (128, 64, 1, False), (3, 128, 8, 8)),
    "rb128_gn1_64x64": ("ResidualBlock", (128, 128, 1, True), (2, 128, 64, 64)),
    "ab64_gn4": ("AttentionBlock", (64, 4, True), (2, 64, 12, 16)),
    "rb32_gn32_simt": ("ResidualBlock", (32, 32, 32, True), (1, 32, 10, 6)),
}


def dense_stride(shape):
    """spatial sampling stride of the stored golden output (full for the small cases)"""
    return 4 if shape[2] * shape[3] > 1024 else 1


def dense_block_inputs(name, cls):
    """(block with deterministic weights, input) for one DENSE_BLOCKS case; cls = the class to instantiate."""
    from mcquic_b200.utils.synthetic import synthetic_block_state
    kind, args, shape = DENSE_BLOCKS[name]
    block = cls(*args).eval()
    block.load_state_dict(synthetic_block_state(block.state_dict(), name, seed=0))
    return block, uniform(shape, name + ".x", 2)


def dense_block_oracle(name, sd, x):
    from oracle import mcquic_oracle as O
    kind, args, _ = DENSE_BLOCKS[name]
    if kind == "AttentionBlock":
        return O.attention_block(_prefixed(sd), "b", x, groups=args[1])
    return O.residual_block(_prefixed(sd), "b", x, groups=args[2])


def _prefixed(sd):
    return {"b." + k: v for k, v in sd.items()}


# `Neon` tokenizer cases (BASELINE configs[4] family at test size): name -> (channel, k, size, denseNorm, n, h, w)
NEON_CASES = {
    "neon_c32_gn": (32, 64, [16, 16, 8, 8], True, 2, 100, 128),     # GroupNorm everywhere, reflect padding, SIMT convs
    "neon_c64_plain": (64, 128, [8, 8], False, 1, 128, 128),        # SiLU blocks, tensor-core trunk (C = 64 / 128)
}


def neon_inputs(name, cls):
    """(model with deterministic weights, image batch) for one NEON_CASES entry; cls = the Neon class to build."""
    from mcquic_b200.utils.synthetic import synthetic_block_state
    c, k, size, dense, n, h, w = NEON_CASES[name]
    model = cls(c, k, list(size), dense).eval()
    model.load_state_dict(synthetic_block_state(model.state_dict(), name, seed=0))
    return model, uniform((n, 3, h, w), name + ".image", 1)


# ---- training step (NEXT-3): tiny models the reference can run forward + backward on CPU in seconds
# name -> (channel, k, size, denseNorm, n, h, w)
TRAIN_CASES = {
    # (`size` = latent grids per level, first = image / 16; it must end in two equal entries: the last level's up-projection
    #  is the identity upstream, quantizer.py:616,641)
    "train_neon_c32_gn": (32, 64, [8, 4, 4], True, 2, 128, 128),     # GroupNorm everywhere, 32 / 64 / 8-channel nets
    "train_neon_c64_plain": (64, 128, [4, 4], False, 2, 64, 64),     # SiLU blocks, 64 / 128-channel trunk
}


class DeterministicRand:
    """context manager: torch.rand_like -> counter-hash uniforms in [0, 1) that depend only on (shape, call number), so the
    reference on CPU and the CUDA path draw the SAME Gumbel noise / drop masks (consume given uniforms
    rather than reproduce Philox offsets)"""

    def __init__(self, tag: str):
        self.tag, self.calls, self._orig = tag, 0, None

    def __enter__(self):
        self._orig = torch.rand_like

        def rand_like(t, *a, **k):
            u = (uniform(tuple(t.shape), f"{self.tag}.rand.{self.calls}", 9) + 1.0) * 0.5
            self.calls += 1
            return u.to(device=t.device, dtype=t.dtype)

        torch.rand_like = rand_like
        return self

    def __exit__(self, *exc):
        torch.rand_like = self._orig


def train_inputs(name, cls):
    from mcquic_b200.utils.synthetic import synthetic_block_state
    c, k, size, dense, n, h, w = TRAIN_CASES[name]
    model = cls(c, k, list(size), dense)
    model.load_state_dict(synthetic_block_state(model.state_dict(), name, seed=0))
    return model, uniform((n, 3, h, w), name + ".image", 1)


def grad_sample(g: torch.Tensor, count: int = 64) -> torch.Tensor:
    """a deterministic strided sample of a gradient tensor (what the golden files store next to its norm)"""
    flat = g.detach().flatten()
    step = max(1, flat.numel() // count)
    return flat[::step][:count].clone()
