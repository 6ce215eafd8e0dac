"""ncu --csv launch list (gpu__time_duration / dram bytes per launch) -> per-kernel summary (text + json)."""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT_T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, steps, out_json=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    by = defaultdict(dict)
    for r in rows[1:]:
        by[r[ix["ID"]]]["name"] = r[ix["Kernel Name"]]
        by[r[ix["ID"]]]["grid"] = r[ix["Grid Size"]]
        by[r[ix["ID"]]][r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for v in by.values():
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", v["name"]))
        t = v["gpu__time_duration.sum"]
        a = agg[name]
        a[0] += 1
        a[1] += t[0] * UNIT_T[t[1]]
        for j, m in ((2, "dram__bytes_read.sum"), (3, "dram__bytes_write.sum")):
            if m in v:
                a[j] += v[m][0] * UNIT_B[v[m][1]]
    total = sum(a[1] for a in agg.values())
    print(f"# {len(by)} launches in {steps} timed step(s); serialised device time {total/1e3/steps:.2f} ms per step")
    print(f"{'kernel':58s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s} {'dram R GB':>10s} {'dram W GB':>10s}")
    out = {"steps": steps, "kernels": {}}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:58]:58s} {a[0]/steps:13.1f} {a[1]/1e3/steps:9.3f} {100*a[1]/total:6.1f}% {a[2]/1e9/steps:10.3f} {a[3]/1e9/steps:10.3f}")
        out["kernels"][name] = {"launches_per_step": a[0] / steps, "ms_per_step": a[1] / 1e3 / steps,
                                "share": a[1] / total, "dram_read_bytes_per_step": a[2] / steps,
                                "dram_write_bytes_per_step": a[3] / steps}
    ours = {k: v for k, v in out["kernels"].items() if k.startswith("mcq::")}
    out["dram_bytes_per_step"] = sum(v["dram_read_bytes_per_step"] + v["dram_write_bytes_per_step"] for v in ours.values())
    conv = {k: v for k, v in ours.items() if "conv_halo" in k or "conv_tc" in k or "conv_pair" in k or "conv_chain" in k}
    out["conv_dram_bytes_per_step"] = sum(v["dram_read_bytes_per_step"] + v["dram_write_bytes_per_step"] for v in conv.values())
    out["conv_share"] = sum(v["share"] for v in conv.values())
    print(f"# tcgen05 conv kernels: {100*out['conv_share']:.1f}% of the step, DRAM traffic {out['conv_dram_bytes_per_step']/1e9:.2f} GB per step")
    # the dominant kernel: the 3-pass CTA-pair kernel in its bulk-store instantiation <PASSES = 3, GN = 0, DRAIN = 2> = exactly
    # the 3x3 stride-1 encode convolutions on >= 32x32 maps (launches with >= 2 tiles per CTA, mcq_api.cu: drain_kind)
    def targs(name):
        inner = name.split("<", 1)[1].rsplit(">", 1)[0] if "<" in name else ""
        return [t.strip().replace("(int)", "").replace("(bool)", "") for t in inner.split(",")]
    dom = [v for k, v in ours.items() if "conv_pair_kernel" in k and targs(k)[:3] in (["3", "0", "2"], ["3", "false", "2"])]
    if not dom:   # launch lists from before the drain variants existed
        dom = [v for k, v in ours.items() if "conv_pair_kernel<3" in k.replace(" ", "") and "true" not in k and ", 1>" not in k]
    if dom:
        n_l = sum(v["launches_per_step"] for v in dom)
        out["dominant_kernel_dram_bytes_per_launch"] = sum(v["dram_read_bytes_per_step"] + v["dram_write_bytes_per_step"]
                                                           for v in dom) / n_l
        out["dominant_kernel_launches_per_step"] = n_l
        out["dominant_kernel_ms_per_step"] = sum(v["ms_per_step"] for v in dom)
        out["dominant_kernel_share"] = sum(v["share"] for v in dom)
        print(f"# dominant kernel conv_pair_kernel<3>: {n_l:.0f} launches/step, {out['dominant_kernel_ms_per_step']:.3f} ms/step "
              f"({100*out['dominant_kernel_share']:.1f}% of the serialised step), DRAM {out['dominant_kernel_dram_bytes_per_launch']/1e6:.1f} MB per launch")
    if out_json:
        json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)
