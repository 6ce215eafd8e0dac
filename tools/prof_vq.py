"""BASELINE configs[2]: qp=3-shaped level-0 VQ alone (M=6 codebooks, K=2048, d=32, batch 32 of 512x512 -> 32x32 grid).
Times mcq_vq_assign with and without the soft logits (CUDA events, L2 flushed between launches) and prints the
achieved algorithmic HBM GB/s / TFLOP/s (x 25.2 MB + codebook 1.57 MB + codes 1.57 MB,
+1.61 GB with logits; 25.77 GFLOP).  Also the ncu driver for that kernel:
   ncu --set full --clock-control none --import-source on -k regex:vq_ -o gpurun_out/vq python tools/prof_vq.py --once
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200.engine import Engine, pack_codebook  # noqa: E402
from mcquic_b200.utils.synthetic import uniform  # noqa: E402

once = "--once" in sys.argv
n, h, w, m, k, d = 32, 32, 32, 6, 2048, 32
if "--qp1" in sys.argv:      # the headline model's level-0 VQ: one codebook over all 128 channels, K = 8192, batch 64 of 256x256
    n, h, w, m, k, d = 64, 16, 16, 1, 8192, 128
eng = Engine()
x = (uniform((n, m * d, h, w), "vq.big", 5) * 0.26).cuda()
cb = (uniform((m, k, d), "vq.bigcb", 5) * 0.19).cuda().contiguous()
xg = eng.from_nchw(x, {"f32"}).f32
c2 = (cb ** 2).sum(-1).contiguous()
packed = pack_codebook(cb)
if "--simt" in sys.argv:
    eng.impl = 1   # _lib.IMPL_SIMT: the FFMA kernel (vq_assign_kernel) for comparison
hist = torch.zeros(m * k, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
P = n * h * w
bytes_hard = P * m * d * 4 + m * k * d * 4 + P * m * 8
bytes_soft = bytes_hard + P * m * k * 4
flops = 2.0 * P * m * k * d
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
out = {}
for logits in (False, True):
    reps = 1 if once else 20
    ts = []
    for it in range(reps + (0 if once else 3)):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = eng.vq_assign(xg, cb, c2, n, h, w, logits=logits, hist=hist, packed=packed)
        e1.record()
        e1.synchronize()
        if once or it >= 3:
            ts.append(e0.elapsed_time(e1))
    del r
    ms = sum(ts) / len(ts)
    b = bytes_soft if logits else bytes_hard
    out["soft" if logits else "hard"] = {"ms": ms, "alg_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peaks["hbm_gbs"],
                                        "TFLOPs": flops / ms / 1e9}
out["shape"] = dict(n=n, h=h, w=w, m=m, k=k, d=d)
print(json.dumps(out))
