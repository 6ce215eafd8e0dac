"""Throughput of the host entropy coder (NEXT-1) next to the reference's own coder (oracle/_ref) on the same codes:
qp=1 code maps of N images, CDFs from a peaked pmf.  CPU only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mcquic_b200 import entropy  # noqa: E402
from oracle import build_ref  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
levels = [(8192, 16), (2048, 8), (512, 4)]
codes, cdfs = [], []
for k, g in levels:
    pmf = rng.random(k) ** 3
    cdfs.append(entropy.pmf_to_quantized_cdf(pmf / pmf.sum())[None])
    codes.append(torch.from_numpy(rng.choice(k, size=(N, 1, g, g), p=pmf / pmf.sum())))
symbols = sum(c.numel() for c in codes)
best = 1e9
for _ in range(5):
    t = time.perf_counter()
    streams = [entropy.encode_level(c, cd) for c, cd in zip(codes, cdfs)]
    best = min(best, time.perf_counter() - t)
print(f"ours   encode: {best*1e3:8.2f} ms  {symbols/best/1e6:8.2f} Msym/s  ({os.cpu_count()} threads)")
bestd = 1e9
for _ in range(5):
    t = time.perf_counter()
    back = [entropy.decode_level(s, 1, g, g, cd) for s, (k, g), cd in zip(streams, levels, cdfs)]
    bestd = min(bestd, time.perf_counter() - t)
assert all(torch.equal(a, b) for a, b in zip(back, codes))
print(f"ours   decode: {bestd*1e3:8.2f} ms  {symbols/bestd/1e6:8.2f} Msym/s")
ref = build_ref.load()
if ref is not None:
    enc, dec = ref.RansEncoder(), ref.RansDecoder()
    t = time.perf_counter()
    rs = []
    for c, cd, (k, g) in zip(codes, cdfs, levels):      # the reference's own loop (entropyCoder.py:113-124)
        lst = [cd[0].tolist()]
        for img in c:
            idx = torch.arange(1)[:, None, None].expand_as(img).flatten().int().tolist()
            rs.append(enc.encodeWithIndexes(img.flatten().int().tolist(), idx, lst, [k + 2], torch.zeros_like(img).flatten().int().tolist()))
    te = time.perf_counter() - t
    t = time.perf_counter()
    i = 0
    for c, cd, (k, g) in zip(codes, cdfs, levels):
        lst = [cd[0].tolist()]
        for img in c:
            idx = [0] * (g * g)
            dec.decodeWithIndexes(rs[i], idx, lst, [k + 2], [0] * (g * g))
            i += 1
    td = time.perf_counter() - t
    flat = [s for lv in streams for s in lv]
    print(f"ref    encode: {te*1e3:8.2f} ms  {symbols/te/1e6:8.2f} Msym/s   decode: {td*1e3:8.2f} ms  {symbols/td/1e6:8.2f} Msym/s   identical streams: {flat == rs}")
