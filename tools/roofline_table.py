"""conv launch list of one eager step (bench.py --dump-profile) -> markdown table per layer shape: launches, time, share,
executed TFLOP/s, fraction of the measured dense peak, and the tensor-pipe time the same work needs at that peak.
   python tools/roofline_table.py profiles/r1e_conv_launch_times.json > profiles/r1e_roofline_by_shape.md
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = json.load(open(sys.argv[1]))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = peaks.get("bf16_tflops_sustained", 1388.6)
convs = [r for r in rows if "other" not in r]
others = [r for r in rows if "other" in r]
agg = collections.OrderedDict()
for r in convs:
    a = agg.setdefault((tuple(r["shape"]), r["passes"]), [0, 0.0, 0.0])
    a[0] += 1; a[1] += r["us"]; a[2] += r["flops"]
tot = sum(v[1] for v in agg.values()) + sum(r["us"] for r in others)
print(f"Per-shape device time of one eager step (every launch bracketed by CUDA events; total {tot / 1e3:.2f} ms; peak = "
      f"{peak:.0f} TFLOP/s measured sustained dense fp16/bf16).\n")
print("| n,h,w,cin,cout,k,stride | passes | launches | ms | share | executed TFLOP/s | of peak | ms at peak |")
print("|---|---|---|---|---|---|---|---|")
for (shape, p), (c, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ex = fl * p / us / 1e6
    print(f"| {','.join(map(str, shape))} | {p} | {c} | {us / 1e3:.3f} | {us / tot:.1%} | {ex:.0f} | {ex / peak:.0%} | {fl * p / peak / 1e9:.3f} |")
oa = collections.OrderedDict()
for r in others:
    a = oa.setdefault(r["other"], [0, 0.0]); a[0] += 1; a[1] += r["us"]
for name, (c, us) in sorted(oa.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | - | {c} | {us / 1e3:.3f} | {us / tot:.1%} | - | - | - |")
