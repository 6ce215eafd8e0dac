"""Experiment: L2-resident sub-batches of the full-resolution layers (Compressor.encode_slice / decode_slice).
   python tools/exp_slices.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcquic_b200 import Compressor
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform

K = [8192, 2048, 512]
model = Compressor(128, 1, K).eval()
model.load_state_dict(synthetic_state_dict(128, 1, K, seed=0))
model = model.cuda()
x = uniform((64, 3, 256, 256), "bench.image.0", 0).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps

ref = model.encode(x)
xr = model.decode(ref)
res = {}
for es in (0, 8, 16, 32):
    model.encode_slice = es
    c = model.encode(x)
    assert all(torch.equal(a, b) for a, b in zip(c, ref)), es
    res[f"encode_slice={es}"] = t(lambda: model.encode(x))
    print(f"encode_slice={es}: {res[f'encode_slice={es}']:.3f} ms", flush=True)
model.encode_slice = 0
for ds in (0, 8, 16, 22, 32):
    model.decode_slice = ds
    y = model.decode(ref)
    assert torch.equal(y, xr), ds
    res[f"decode_slice={ds}"] = t(lambda: model.decode(ref))
    print(f"decode_slice={ds}: {res[f'decode_slice={ds}']:.3f} ms", flush=True)
print(json.dumps(res))
