"""Experiment: does running the two halves of a batch as two independent graph replays on two streams hide the
latency-bound small-map phases of one half behind the full-resolution layers of the other?  (qp=1, 64 x 256x256)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200 import Compressor  # noqa: E402
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402

K = [8192, 2048, 512]
sd = synthetic_state_dict(128, 1, K, seed=0)


def make():
    m = Compressor(128, 1, K).eval()
    m.load_state_dict(sd)
    return m.cuda()


x = uniform((64, 3, 256, 256), "bench.image.0", 0).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
full = make()


def step_full():
    return full.decode(full.encode(x))


parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
models = [make() for _ in range(parts)]            # separate engines / graphs / side streams per part
streams = [torch.cuda.Stream() for _ in range(parts)]
xs = list(x.chunk(parts))
offset_decode = len(sys.argv) > 2 and sys.argv[2] == "stagger"


def step_parts():
    cur = torch.cuda.current_stream()
    outs = [None] * parts
    for i, (m, s) in enumerate(zip(models, streams)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs[i] = m.decode(m.encode(xs[i]))
    for s in streams:
        cur.wait_stream(s)
    return outs


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


a = timed(step_full)
b = timed(step_parts)
ref = step_full()
got = torch.cat(step_parts())
print({"full_ms(median,min)": a, f"{parts}_parts_ms": b, "identical": bool(torch.equal(ref, got))})
