"""Small driver for ncu: a few launches of the big-layer convolution shapes (N=64).  Usage (on the GPU box):
   ncu --set full --clock-control none --import-source on -k regex:conv_ -s 6 -c 4 -o gpurun_out/prof python tools/prof_conv.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from convcase import make_planes  # noqa: E402
from mcquic_b200 import _lib  # noqa: E402
from mcquic_b200.engine import Act, Engine, pack_conv  # noqa: E402

_lib.apply_options(os.environ.get("MCQ_OPTIONS", ""))   # e.g. MCQ_OPTIONS=direct_epi=1 (applied here through mcq_set_option)

eng = Engine("tcgen05")
g = torch.Generator().manual_seed(0)
n, h, w, c = 64, int(os.environ.get("PROF_HW", "128")), int(os.environ.get("PROF_HW", "128")), 128
x = torch.randn(n, h, w, c, generator=g).cuda()
res = torch.randn(n, h, w, c, generator=g).cuda()
wt = ((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
pc = pack_conv(wt, torch.zeros(c).cuda(), 1, 0, "cuda")
for passes in (3, 1):
    eng.passes = passes
    a = make_planes(x, passes)
    for _ in range(3):   # 3 launches per variant: ncu -s 2 -c 1 etc.
        eng.conv(pc, a, Act(n, h, w, c), {"f32", "silu"}, res1=res)
torch.cuda.synchronize()
print("done")
