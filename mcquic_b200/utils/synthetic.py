"""Deterministic synthetic weights in the reference's state_dict layout.

There is no network here, so no pretrained checkpoint (`qp_2_msssim`, mcquic/demo.py:20-24) can be fetched:
benchmarks, tests and golden vectors use weights produced by a counter-based integer hash (splitmix64), which
gives bit-identical tensors on every machine and numpy/torch version -- unlike seeded torch initialisers.
Magnitudes follow the reference's initialisers: conv weights/biases U(+-1/sqrt(fan_in)) (nn.Conv2d default),
codebooks with the "SmallInit" std sqrt(2/(5d)) (mcquic/modules/quantizer.py:398), GDN beta/gamma at their
reference init (mcquic/nn/gdn.py:53-63) plus a small perturbation so the off-diagonal path is exercised.
"""
import zlib
from typing import Dict, List

import numpy as np
import torch

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return x ^ (x >> np.uint64(31))


def uniform(shape, key: str, seed: int) -> torch.Tensor:
    """U[-1, 1) fp32 tensor that depends only on (shape, key, seed)."""
    count = int(np.prod(shape))
    base = np.uint64((zlib.crc32(key.encode()) << 32) ^ (seed * 0x632BE5AB + 0x1234567))
    with np.errstate(over="ignore"):
        h = _splitmix64(np.arange(count, dtype=np.uint64) + base)
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)  # 24 random bits -> [0, 1)
    return torch.from_numpy((u * 2.0 - 1.0).astype(np.float32).reshape(shape))


def synthetic_state_dict(channel: int, m: int, k: List[int], seed: int = 0) -> Dict[str, torch.Tensor]:
    from ..modules.compressor import Compressor
    with torch.no_grad():
        torch_state = torch.get_rng_state()
        template = Compressor(channel, m, list(k)).state_dict()
        torch.set_rng_state(torch_state)
    out: Dict[str, torch.Tensor] = {}
    codebooks: Dict[str, torch.Tensor] = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        if key.endswith("._codebook"):
            # the three aliased keys of a level share one tensor
            level = key.split("._encoders.")[-1].split("._decoders.")[-1].split(".")[0]
            if level not in codebooks:
                std = (2.0 / (5.0 * shape[-1])) ** 0.5
                codebooks[level] = uniform(shape, f"codebook.{level}", seed) * (std * 3.0 ** 0.5)
            out[key] = codebooks[level].clone()
        elif key.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out[key] = uniform(shape, key, seed) / fan_in ** 0.5
        elif key.endswith(".bias"):
            w = template[key[:-4] + "weight"]
            fan_in = w.shape[1] * w.shape[2] * w.shape[3]
            out[key] = uniform(shape, key, seed) / fan_in ** 0.5
        elif key.endswith(".gamma"):
            out[key] = ref.clone() + uniform(shape, key, seed).abs() * 0.02
        elif key.endswith(".beta"):
            out[key] = ref.clone() + uniform(shape, key, seed).abs() * 0.1
        else:  # reparam constants, temperatures, freqEMA, bounds: keep the reference's init values
            out[key] = ref.clone()
    return out


def synthetic_block_state(template: Dict[str, torch.Tensor], tag: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights for ONE block (`template` = its state_dict): convolutions as above, nn.GroupNorm
    affine parameters (1-D `weight` / `bias` without a 4-D sibling) as 1 + 0.2 u and 0.1 u."""
    out: Dict[str, torch.Tensor] = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        name = f"{tag}.{key}"
        if key.endswith("._codebook"):
            # every codebook key of a ResidualBackwardQuantizer aliases ONE [1, k, d] tensor (quantizer.py:596)
            std = (2.0 / (5.0 * shape[-1])) ** 0.5
            out[key] = uniform(shape, f"{tag}.codebook", seed) * (std * 3.0 ** 0.5)
        elif key.endswith(".weight") and len(shape) == 4:
            out[key] = uniform(shape, name, seed) / (shape[1] * shape[2] * shape[3]) ** 0.5
        elif key.endswith(".gamma"):
            out[key] = ref.clone() + uniform(shape, name, seed).abs() * 0.02
        elif key.endswith(".beta"):
            out[key] = ref.clone() + uniform(shape, name, seed).abs() * 0.1
        elif key.endswith(".bias") and template[key[:-4] + "weight"].dim() == 4:
            w = template[key[:-4] + "weight"]
            out[key] = uniform(shape, name, seed) / (w.shape[1] * w.shape[2] * w.shape[3]) ** 0.5
        elif key.endswith(".weight") and len(shape) == 1:
            out[key] = 1.0 + 0.2 * uniform(shape, name, seed)
        elif key.endswith(".bias"):
            out[key] = 0.1 * uniform(shape, name, seed)
        else:
            out[key] = ref.clone()
    return out
