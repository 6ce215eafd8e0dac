"""In-situ latency of dependent convolution launches: a chain of L identical 3x3 convs (layer i reads layer i-1's
planes), captured as ONE CUDA graph on one stream and replayed.  time / L = what one more layer costs inside
encode/decode (kernel + dependent-launch gap), without the CPU launch cost an eager loop adds and without the
cold-cache serialisation of an ncu launch list.  Also: two such chains on two streams (the AttentionBlock pattern)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from convcase import make_planes  # noqa: E402
from mcquic_b200.engine import Act, Engine, pack_conv  # noqa: E402

L = 24
n, c = 64, 128
eng = Engine("tcgen05")
eng.chain = False
g = torch.Generator().manual_seed(0)
packs = [pack_conv(((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda(), torch.zeros(c).cuda(), 1, 0, "cuda")
         for _ in range(2 * L)]
out = {}
for hw in ([int(v) for v in os.environ["PROF_HW"].split(",")] if os.environ.get("PROF_HW") else (4, 8, 16, 32, 64)):
    for passes in (1, 3):
        eng.passes = passes
        x = torch.randn(n, hw, hw, c, generator=g).cuda()
        a0 = make_planes(x, passes)

        def chain(off):
            a, act = a0, Act(n, hw, hw, c)
            for i in range(L):
                o = eng.conv(packs[off + i], a, act, {"silu"})
                a, act = o.silu, o
            return a

        def one():
            return chain(0)

        def two():
            return eng.parallel(lambda: chain(0), lambda: chain(L))

        for name, body in (("one_stream", one), ("two_streams", two)):
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                body()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = body()
            ts = []
            for it in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            us = 1e3 * min(ts) / L
            flops = 2.0 * n * hw * hw * c * c * 9 * passes
            out[f"{hw}x{hw} p{passes} {name}"] = round(us, 2)
            print(f"{hw:3d}x{hw:<3d} passes {passes} {name:12s}: {us:7.2f} us per layer"
                  f" (per branch-layer {us / (2 if name == 'two_streams' else 1):6.2f}; MMA-only floor {flops / 1388e12 * 1e6:6.2f} us)", flush=True)
            del gr, keep
print(json.dumps(out))
