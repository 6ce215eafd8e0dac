"""Experiment: one batch of 64 as 1 x 64 vs 2 x 32 / 4 x 16 on concurrent streams (graph replays)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mcquic_b200 import Compressor  # noqa: E402
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402

K = [8192, 2048, 512]
sd = synthetic_state_dict(128, 1, K, seed=0)
x = uniform((64, 3, 256, 256), "bench.image.0", 0).cuda()
for parts in (1, 2, 4):
    models = []
    for _ in range(parts):
        m = Compressor(128, 1, K).eval()
        m.load_state_dict(sd)
        models.append(m.cuda())
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = x.chunk(parts)

    def step():
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for m, st, c in zip(models, streams, chunks):
            with torch.cuda.stream(st):
                m.decode(m.encode(c))
        for st in streams:
            cur.wait_stream(st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{parts} x {64 // parts}: {ms:.2f} ms/step = {64 * 65536 / ms / 1e3:.1f} MPix/s", flush=True)
