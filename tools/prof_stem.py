"""Timing of the stem layer (64 x 3 x 256 x 256 -> 128 x 128 x 128): tcgen05 kernel vs FFMA kernel, fp32 / uint8 input,
with and without stores (epi_skip).   python tools/prof_stem.py [--once]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcquic_b200 import _lib
from mcquic_b200.engine import Engine
from mcquic_b200.nn import conv3x3
from mcquic_b200.utils.synthetic import uniform

conv = conv3x3(3, 128, 2).cuda()
x = uniform((64, 3, 256, 256), "bench.image.0", 0).cuda()
xu = ((x + 1) * 127.5).round().clamp(0, 255).to(torch.uint8)
eng = Engine()
eng.passes = 3
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
once = "--once" in sys.argv


def t(fn, reps=10):
    for _ in range(2):
        fn()
    if once:
        return 0.0
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3

res = {}
for name, tc, inp, skip in (("tc fp32", True, x, 0), ("tc uint8", True, xu, 0), ("tc fp32, no stores", True, x, 1),
                            ("ffma fp32", False, x, 0)):
    if once and name != "tc fp32":
        continue
    eng.stem_tc = tc
    _lib.set_option("epi_skip", skip)
    res[name] = t(lambda: eng.stem(conv, inp, (0, 0, 256, 256), {"f32", "silu"}))
    print(f"{name}: {res[name]:.1f} us", flush=True)
_lib.set_option("epi_skip", 0)
print(json.dumps(res))
