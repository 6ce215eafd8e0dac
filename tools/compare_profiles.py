"""Per (shape, passes, epilogue kind) device time of two `bench.py --dump-profile` files side by side (A/B of a library
option over every launch class of the step).  Usage: python tools/compare_profiles.py a.json b.json"""
import json
import sys
from collections import defaultdict


def load(path):
    acc = defaultdict(lambda: [0, 0.0])
    for e in json.load(open(path)):
        key = ("other", e["other"]) if "other" in e else (tuple(e["shape"]), e["passes"], e.get("epi"))
        acc[key][0] += 1
        acc[key][1] += e["us"]
    return acc


a, b = load(sys.argv[1]), load(sys.argv[2])
rows = sorted(set(a) | set(b), key=lambda k: -(a.get(k, [0, 0])[1]))
ta = tb = 0.0
print(f"{'launch class':58s} {'n':>4s} {'A us':>9s} {'B us':>9s} {'B-A':>8s}")
for k in rows:
    na, ua = a.get(k, [0, 0.0])
    nb, ub = b.get(k, [0, 0.0])
    ta += ua
    tb += ub
    print(f"{str(k):58s} {na:4d} {ua:9.1f} {ub:9.1f} {ub - ua:+8.1f}")
print(f"{'total':58s} {'':4s} {ta:9.1f} {tb:9.1f} {tb - ta:+8.1f}")
