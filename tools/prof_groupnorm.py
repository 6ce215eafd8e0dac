"""Timing + ncu driver for mcq_groupnorm at the shapes a dense-norm ResidualBlock sees at batch 64 (C=128, groups 32):
   python tools/prof_groupnorm.py            # CUDA-event timing, L2 flushed between launches -> JSON line
   ncu --set full --clock-control none --import-source on -k regex:groupnorm -s 2 -c 1 -o gpurun_out/x/gn python tools/prof_groupnorm.py
Algorithmic bytes per launch: fp32 activation in (4 B/elem) + the split-fp16 plane pair out (2 x 2 B/elem; hi only: 2 B).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200.engine import Act, Engine  # noqa: E402

eng = Engine("tcgen05")
res = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
for n, hw, c, groups, passes in ((64, 64, 128, 32, 3), (64, 128, 128, 32, 3), (64, 64, 128, 32, 1), (64, 16, 128, 32, 3),
                                 (64, 64, 256, 32, 3)):
    eng.passes = passes
    x = torch.randn(n, hw, hw, c, device="cuda")
    norm = torch.nn.GroupNorm(groups, c).cuda()
    act = Act(n, hw, hw, c, f32=x)
    for _ in range(3):
        eng.groupnorm(norm, act, {"raw"})
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.groupnorm(norm, act, {"raw"})
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    nbytes = x.numel() * (4 + (4 if passes == 3 else 2))
    res.append({"n": n, "hw": hw, "c": c, "groups": groups, "passes": passes, "ms": ms, "alg_bytes": nbytes,
                "GBps": nbytes / ms / 1e6, "frac_of_measured_hbm": (nbytes / ms / 1e6) / peaks["hbm_gbs"] if "hbm_gbs" in peaks else None})
print(json.dumps({"groupnorm": res, "hbm_peak_gbs": peaks.get("hbm_gbs")}))
