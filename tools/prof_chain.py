"""Micro-benchmark of the layer-chain kernel: a stack of ResidualBlocks (2 convs each) on a small map, issued as
one chain vs layer by layer.  Usage (GPU box):
    python tools/prof_chain.py                 # timing table
    PROF_HW=4 PROF_PASSES=1 PROF_ONLY=chain ncu --set full ... python tools/prof_chain.py
Env knobs read by the library: MCQ_CHAIN_IPC, MCQ_CHAIN_NOSYNC (timing only, wrong results), MCQ_EPI_SKIP.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from convcase import make_planes  # noqa: E402
from mcquic_b200.engine import Act, Engine  # noqa: E402
from mcquic_b200.nn import ResidualBlock  # noqa: E402

N = int(os.environ.get("PROF_N", "64"))
BLOCKS = int(os.environ.get("PROF_BLOCKS", "6"))
only = os.environ.get("PROF_ONLY", "")
sizes = [int(os.environ["PROF_HW"])] if "PROF_HW" in os.environ else [4, 8, 16]
passes_list = [int(os.environ["PROF_PASSES"])] if "PROF_PASSES" in os.environ else [1, 3]

torch.manual_seed(0)
eng = Engine("tcgen05")
eng.multistream = False
blocks = [ResidualBlock(128, 128).cuda() for _ in range(BLOCKS)]


WARM = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)


def run(hw, passes, chain, reps):
    """GPU time of the launches themselves (library-side events), queued behind ~20 ms of matmuls so that the clocks
    are up and no host time leaks into the measurement."""
    eng.passes = passes
    eng.chain = chain
    x = torch.randn(N, hw, hw, 128, device="cuda") * 0.5
    act = Act(N, hw, hw, 128, f32=x, silu=make_planes(torch.nn.functional.silu(x), passes))
    best = float("inf")
    for r in range(reps + 1):
        for _ in range(12):
            WARM @ WARM
        eng.profile = []
        eng.run_seq(blocks, act, {"f32", "silu"})
        eng.flush()
        torch.cuda.synchronize()
        us = sum(p["ev"][0].elapsed_time(p["ev"][1]) for p in eng.profile) * 1e3
        eng.profile = None
        if r > 0:
            best = min(best, us)
    return best


for hw in sizes:
    for passes in passes_list:
        row = []
        for chain in (True, False):
            if only and only != ("chain" if chain else "single"):
                continue
            us = run(hw, passes, chain, 3)
            row.append(f"{'chain' if chain else 'single'} {us:8.1f} us = {us / (2 * BLOCKS):6.2f} us/layer")
        print(f"n{N} {hw}x{hw} passes={passes} layers={2 * BLOCKS}: " + " | ".join(row), flush=True)

if os.environ.get("PROF_TIMELINE"):
    hw, passes = sizes[0], passes_list[0]
    buf = torch.zeros(3072 + 512, dtype=torch.int64, device="cuda")
    eng.lib.mcq_debug_timeline(buf.data_ptr())
    run(hw, passes, True, 1)
    torch.cuda.synchronize()
    eng.lib.mcq_debug_timeline(None)
    t = buf.cpu().tolist()
    prod = sorted(v for v in t[:1024] if v)
    mma = [v for v in t[1024:2048] if v]
    epi = t[2048:2560]
    sync = t[2560:3072]
    multi = t[3072:]
    t0 = min(prod[0], mma[0])
    print("producer issue times (cycles):", [v - t0 for v in prod[:60]])
    print("mma stage-arrival times     :", [v - t0 for v in mma[:60]])
    print("epilogue (begin,end) per tile:", [(epi[2 * i] - t0, epi[2 * i + 1] - t0) for i in range(6) if epi[2 * i]])
    for l in range(1, 5):
        e = sync[8 * l:8 * l + 8]
        print(f"layer {l}: epi fence-begin {e[0]-t0} fence-end {e[1]-t0} barrier-done {e[2]-t0} | producer at-barrier {e[4]-t0} "
              f"barrier-done {e[5]-t0} fence-done {e[6]-t0}")
    g0 = min(multi[64 * c + 4] for c in range(8))
    for l in (1, 2, 3):
        print(f"layer {l} (globaltimer ns): " + " ".join(
            f"cta{c}: arrive {multi[64*c+4*l]-g0} done {multi[64*c+4*l+2]-g0} (clk {multi[64*c+4*l+3]-multi[64*c+4*l+1]})" for c in range(8)))
