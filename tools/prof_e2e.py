"""Where the end-to-end step (host buffers) spends time beyond the device-resident step: encode / decode, device vs
host I/O variants, CUDA events on the main stream (L2 not flushed), plus raw PCIe copy times of the same buffers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200 import Compressor  # noqa: E402
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402

C, M, K = 128, 1, [8192, 2048, 512]
model = Compressor(C, M, K).eval()
model.load_state_dict(synthetic_state_dict(C, M, K, seed=0))
model = model.cuda()
xh = uniform((64, 3, 256, 256), "bench.image.0", 0).pin_memory()
xd = xh.cuda()
oh = torch.empty((64, 3, 256, 256)).pin_memory()


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        t = e0.elapsed_time(e1)
        best, tot = min(best, t), tot + t
    return best, tot / reps


codes = model.encode(xd)
rows = [
    ("H2D 50 MB (pinned -> device)", lambda: xd.copy_(xh, non_blocking=True)),
    ("D2H 50 MB (device -> pinned)", lambda: oh.copy_(xd, non_blocking=True)),
    ("encode(device)", lambda: model.encode(xd)),
    ("encode(pinned host)", lambda: model.encode(xh)),
    ("decode(codes)", lambda: model.decode(codes)),
    ("decode(codes, out=pinned host)", lambda: model.decode(codes, out=oh)),
]
for name, fn in rows:
    b, a = timed(fn)
    print(f"{name:34s} best {b:8.3f} ms   mean {a:8.3f} ms", flush=True)
