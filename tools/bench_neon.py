"""Inference timing of the `Neon` tokenizer at the shapes of configs/a800_16.yaml (channel 256, k 4096, 17 levels; the
only Neon config that instantiates at the reference's HEAD, SURVEY.md finding 3): encode + decode of n 512x512 images,
CUDA graphs, CUDA events, synthetic weights.  Not the BASELINE metric (that is bench.py); reported next to it as the
first measurement of the configs[4] model family on this path.
   python tools/bench_neon.py [--n 8] [--hw 512] [--channel 256] [--steps 5]
Algorithmic FLOPs: 2*MACs of every Conv2d (hook-counted by SURVEY.md 8a row a16: 12.3 TFLOP per 512x512 image, fwd).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200 import Neon, _lib  # noqa: E402
from mcquic_b200.utils.synthetic import synthetic_block_state, uniform  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--hw", type=int, default=512)
ap.add_argument("--channel", type=int, default=256)
ap.add_argument("--k", type=int, default=4096)
ap.add_argument("--dense", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--opt", default="", help="library options name=value,... (mcq_set_option), e.g. direct_epi=1")
ap.add_argument("--decode-passes", type=int, default=3)
ap.add_argument("--layers", type=int, default=0, help="also print the N most expensive conv shapes of the eager pass")
args = ap.parse_args()
from mcquic_b200 import _lib as _mcq_lib  # noqa: E402
_mcq_lib.apply_options(args.opt)
size = [16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1]
if os.environ.get("NEON_SIZE"):
    size = [int(v) for v in os.environ["NEON_SIZE"].split(",")]

t0 = time.time()
model = Neon(args.channel, args.k, size, bool(args.dense)).eval()
model.load_state_dict(synthetic_block_state(model.state_dict(), "neon.bench", seed=0))
model = model.cuda()
model.decode_passes = args.decode_passes
params = sum(p.numel() for p in set(model.parameters()))
x = uniform((args.n, 3, args.hw, args.hw), "neon.bench.image", 0).cuda()
conv_flops = 0.0
eng = model.engine
eng.profile = []
model.use_graphs = False
codes = model.encode(x)
xhat = model.decode(codes)
torch.cuda.synchronize()
eng.profile = [r for r in eng.profile if "other" not in r]     # convolution launches only
flops = sum(r["flops"] for r in eng.profile)
launches_eager = len(eng.profile)
layers = {}
for r in eng.profile:
    key = (tuple(r["shape"]), r["passes"], "simt" if r["impl"] == _lib.IMPL_SIMT else "tc")
    ms = r["ev"][0].elapsed_time(r["ev"][1])
    cur = layers.setdefault(key, [0, 0.0, 0.0])
    cur[0] += 1; cur[1] += ms; cur[2] += r["flops"]
eng.profile = None
model.use_graphs = True
setup_s = time.time() - t0
for _ in range(2):
    codes = model.encode(x)
    xhat = model.decode(codes)
torch.cuda.synchronize()
enc, dec = [], []
for _ in range(args.steps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    codes = model.encode(x)
    e1.record()
    xhat = model.decode(codes)
    e2.record()
    torch.cuda.synchronize()
    enc.append(e0.elapsed_time(e1))
    dec.append(e1.elapsed_time(e2))
enc.sort(); dec.sort()
e, d = enc[len(enc) // 2], dec[len(dec) // 2]
tokens = sum(int(c[0].numel()) for c in codes)
print(json.dumps({
    "model": f"Neon({args.channel}, {args.k}, {size if len(size) != 17 else 'size17'}, denseNorm={bool(args.dense)})", "params_M": params / 1e6,
    "n": args.n, "hw": args.hw, "encode_ms": e, "decode_ms": d, "images_per_s": args.n / (e + d) * 1e3,
    "tokens_per_image": tokens, "tokens_per_s": args.n * tokens / (e + d) * 1e3,
    "conv_tflop_per_step": flops / 1e12, "alg_tflops": flops / (e + d) / 1e9, "conv_launches": launches_eager,
    "passes": {"encode": 3, "decode": args.decode_passes}, "finite": bool(torch.isfinite(xhat).all()),
    "device_error_flag": int(_lib.load().mcq_device_error_flag()), "setup_s": setup_s,
    "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))
if args.layers:
    tot = sum(v[1] for v in layers.values())
    print(f"# eager conv time {tot:.1f} ms over {launches_eager} launches; (n,h,w,cin,cout,k,stride) passes impl: count, ms, share, executed TFLOP/s")
    for key, (cnt, ms, fl) in sorted(layers.items(), key=lambda kv: -kv[1][1])[:args.layers]:
        print(f"# {key[0]} p{key[1]} {key[2]}: {cnt:4d} {ms:8.2f} ms {ms / tot:6.1%} {fl * key[1] / ms / 1e9:8.1f}")
