"""ncu driver for the HBM-bound side kernels at the benchmark shapes (N=64): the inverse-GDN 1x1 convolution at
128x128 (1-pass), the GDN 1x1 at 64x64 (3-pass), the attention gate 1x1 at 64x64 and the RGB stem.
   ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem" -o gpurun_out/x/misc python tools/prof_misc.py
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

eng = Engine("tcgen05")
g = torch.Generator().manual_seed(0)
n, c = 64, 128
gamma = (torch.rand(c, c, 1, 1, generator=g) * 0.1 + 0.01).cuda()
beta = (torch.rand(c, generator=g) + 0.5).cuda()
pc = pack_conv(gamma, beta, 1, 0, "cuda")
wg = ((torch.rand(c, c, 1, 1, generator=g) * 2 - 1) / c ** 0.5).cuda()
pg = pack_conv(wg, torch.zeros(c).cuda(), 1, 0, "cuda")
for hw, passes, mode in ((128, 1, _lib.EPI_IGDN), (64, 3, _lib.EPI_GDN), (64, 1, _lib.EPI_GATE)):
    eng.passes = passes
    u = torch.randn(n, hw, hw, c, generator=g).cuda()
    sq = make_planes(u * u, passes)
    if mode == _lib.EPI_GATE:
        r = torch.randn(n, hw, hw, c, generator=g).cuda()
        eng.conv(pg, make_planes(u, passes), Act(n, hw, hw, c), {"f32", "silu"}, mode=mode, res1=r, aux=u)
    else:
        eng.conv(pc, sq, Act(n, hw, hw, c), {"raw"}, mode=mode, aux=u)
    torch.cuda.synchronize()
conv = torch.nn.Conv2d(3, c, 3, 2, 1).cuda()
x = (torch.rand(n, 3, 256, 256, generator=g) * 2 - 1).cuda()
eng.passes = 3
eng.stem(conv, x, (0, 0, 256, 256), {"f32", "silu"})
torch.cuda.synchronize()
print("done")
