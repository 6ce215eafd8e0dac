"""The dense-norm ResidualBlock's first half at batch 64 (conv3x3 128 -> 128 on 64x64 maps, then GroupNorm(32)):
(a) conv + stand-alone mcq_groupnorm (statistics pass re-reads the activation), (b) conv whose epilogue emits the
GroupNorm partials + mcq_groupnorm_apply (one streaming pass).  CUDA events, L2 flushed between launches.
   python tools/prof_gn_fused.py
   ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 2 -c 1 -o gpurun_out/x/gn_apply python tools/prof_gn_fused.py
"""
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

eng = Engine("tcgen05")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


res = []
g = torch.Generator().manual_seed(0)
for n, hw, c, groups, passes in ((64, 64, 128, 32, 3), (64, 64, 128, 32, 1), (64, 128, 128, 32, 3), (16, 64, 256, 32, 3)):
    eng.passes = passes
    x = torch.randn(n, hw, hw, c, generator=g).cuda()
    wt = ((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
    pc = pack_conv(wt, torch.zeros(c).cuda(), 1, _lib.STORE_NHWC, "cuda")
    a = make_planes(x, passes)
    norm = torch.nn.GroupNorm(groups, c).cuda()
    act = Act(n, hw, hw, c)
    plain = eng.conv(pc, a, act, {"f32"})
    fused = eng.conv(pc, a, act, {"f32"}, gn_groups=groups)
    assert fused.gn is not None
    for _ in range(2):
        eng.groupnorm(norm, plain, {"raw"})
        eng.groupnorm(norm, fused, {"raw"})
    r = {"n": n, "hw": hw, "c": c, "groups": groups, "passes": passes,
         "conv_ms": timed(lambda: eng.conv(pc, a, act, {"f32"})),
         "conv_with_stats_ms": timed(lambda: eng.conv(pc, a, act, {"f32"}, gn_groups=groups)),
         "groupnorm_standalone_ms": timed(lambda: eng.groupnorm(norm, plain, {"raw"})),
         "groupnorm_apply_ms": timed(lambda: eng.groupnorm(norm, fused, {"raw"}))}
    nbytes = x.numel() * (4 + (4 if passes == 3 else 2))
    r["apply_alg_bytes"] = nbytes
    r["apply_GBps"] = nbytes / r["groupnorm_apply_ms"] / 1e6
    r["standalone_GBps"] = nbytes / r["groupnorm_standalone_ms"] / 1e6
    if "hbm_gbs" in peaks:
        r["apply_frac_of_measured_hbm"] = r["apply_GBps"] / peaks["hbm_gbs"]
    r["pair_ms_before"] = r["conv_ms"] + r["groupnorm_standalone_ms"]
    r["pair_ms_fused"] = r["conv_with_stats_ms"] + r["groupnorm_apply_ms"]
    res.append(r)
print(json.dumps({"gn_fused": res, "hbm_peak_gbs": peaks.get("hbm_gbs")}))
