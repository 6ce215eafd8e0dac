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
