"""In-situ cost of the drain variants: a chain of L dependent 3x3 convolutions captured as one CUDA graph, per layer time
for each (map size, passes, epilogue kind, direct_epi option).

  plane : plane -> plane ({"silu"} output only: first conv of a ResidualBlock)
  full  : plane -> fp32 + SiLU planes with an fp32 residual read (second conv of a ResidualBlock)

Usage (GPU box):  python tools/prof_drain.py [hw,hw,...] [opt,opt,...]   -> one JSON line; direct_epi values A/B'd (default
1 = row per lane with global stores, 5 = bulk-tensor stores wherever they apply)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from convcase import make_planes  # noqa: E402
from mcquic_b200 import _lib  # noqa: E402
from mcquic_b200.engine import Act, Engine, pack_conv  # noqa: E402

L = 12
c = 128
eng = Engine("tcgen05")
eng.chain = False
g = torch.Generator().manual_seed(0)
packs = [pack_conv(((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda(), torch.zeros(c).cuda(), 1, 0, "cuda")
         for _ in range(L)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 128, 16]
OPTS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 5]
_lib.apply_options(os.environ.get("MCQ_OPTIONS", ""))   # further A/B knobs, e.g. MCQ_OPTIONS=pair_nbs=3
out = {}
for hw in sizes:
    n = 64 if hw <= 64 else 32
    for passes in (3, 1):
        eng.passes = passes
        x = torch.randn(n, hw, hw, c, generator=g).cuda() * 0.5
        a0 = make_planes(x, passes)
        for kind in ("plane", "full"):
            for opt in OPTS:
                _lib.set_option("direct_epi", opt)

                def body():
                    a, act, res = a0, Act(n, hw, hw, c), x
                    for i in range(L):
                        if kind == "plane":
                            o = eng.conv(packs[i], a, act, {"silu"})
                        else:
                            o = eng.conv(packs[i], a, act, {"f32", "silu"}, res1=res, res1_scale=0.5)
                            res = o.f32
                        a, act = o.silu, o
                    return a

                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    body()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    keep = body()
                ts = []
                for it in range(6):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gr.replay()
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                us = 1e3 * min(ts) / L
                tf = 2.0 * n * hw * hw * c * c * 9 * passes / us / 1e6
                key = f"{n}x{hw}x{hw} p{passes} {kind} direct_epi={opt}"
                out[key] = {"us_per_layer": round(us, 2), "executed_TFLOPs": round(tf, 1)}
                print(f"{key:44s}: {us:8.2f} us per layer  {tf:7.1f} TFLOP/s executed", flush=True)
                del gr, keep
_lib.set_option("direct_epi", 3)
print(json.dumps(out))
