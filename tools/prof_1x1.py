"""In-situ time of the 1x1 epilogue-heavy layers (GDN / IGDN / gate) on big maps: a run of L independent launches in one
CUDA graph, with and without the drain's loads / stores (epi_skip).  python tools/prof_1x1.py"""
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

_lib.apply_options(os.environ.get("MCQ_OPTIONS", ""))
L, c = 8, 128
eng = Engine("tcgen05")
eng.chain = False
g = torch.Generator().manual_seed(0)
wt = (torch.rand(c, c, 1, 1, generator=g).abs() + 0.01).cuda()
pc = pack_conv(wt, torch.ones(c).cuda(), 1, 0, "cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
out = {}
for (n, hw, passes) in ((32, 128, 1), (32, 64, 1), (64, 64, 3)):
    eng.passes = passes
    xs = [torch.rand(n, hw, hw, c, generator=g).cuda() for _ in range(2)]
    planes = [make_planes(x * x, passes) for x in xs]
    for mode, name in ((_lib.EPI_IGDN, "igdn"), (_lib.EPI_GATE, "gate")):
        for skip in (0, 1):
            _lib.set_option("epi_skip", skip)

            def body():
                outs = []
                for i in range(L):
                    if mode == _lib.EPI_GATE:
                        outs.append(eng.conv(pc, planes[i % 2], Act(n, hw, hw, c), {"f32", "silu"}, mode=mode, res1=xs[0],
                                             aux=xs[1]))
                    else:
                        outs.append(eng.conv(pc, planes[i % 2], Act(n, hw, hw, c), {"raw"}, mode=mode, aux=xs[i % 2]))
                return outs

            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                body()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = body()
            ts = []
            for _ in range(5):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            us = 1e3 * min(ts) / L
            elems = n * hw * hw * c
            if mode == _lib.EPI_GATE:
                nbytes = elems * ((2 if passes == 1 else 4) + 8 + 4 + (2 if passes == 1 else 4))
            else:
                nbytes = elems * ((2 if passes == 1 else 4) + 4 + (2 if passes == 1 else 4))
            key = f"{n}x{hw}x{hw} p{passes} {name} epi_skip={skip}"
            out[key] = {"us": round(us, 1), "alg_TBps": round(nbytes / us / 1e6, 2)}
            print(f"{key:40s}: {us:8.1f} us   {nbytes / us / 1e6:5.2f} TB/s algorithmic", flush=True)
            del gr, keep
_lib.set_option("epi_skip", 0)
print(json.dumps(out))
