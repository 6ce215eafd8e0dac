"""encode + decode timing of any `Compressor(channel, m, k)` shape (CUDA graphs, CUDA events, synthetic weights) with the
per-shape conv profile of one eager pass -- e.g. BASELINE configs[2]'s whole model, not only its VQ kernel:
   python tools/bench_compressor.py --channel 192 --m 6 --k 2048,2048,2048 --n 32 --hw 512 --layers 12
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mcquic_b200 import Compressor, _lib  # noqa: E402
from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--channel", type=int, default=192)
ap.add_argument("--m", type=int, default=6)
ap.add_argument("--k", type=str, default="2048,2048,2048")
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--hw", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--opt", default="", help="library options name=value,... (mcq_set_option), e.g. direct_epi=1")
ap.add_argument("--layers", type=int, default=0)
args = ap.parse_args()
from mcquic_b200 import _lib as _mcq_lib  # noqa: E402
_mcq_lib.apply_options(args.opt)
k = [int(v) for v in args.k.split(",")]
model = Compressor(args.channel, args.m, k).eval()
model.load_state_dict(synthetic_state_dict(args.channel, args.m, k, seed=0))
model = model.cuda()
x = uniform((args.n, 3, args.hw, args.hw), "bench.compressor.image", 0).cuda()
eng = model.engine
eng.profile = []
model.use_graphs = False
codes = model.encode(x)
xhat = model.decode(codes)
torch.cuda.synchronize()
eng.profile = [r for r in eng.profile if "other" not in r]     # convolution launches only
flops = sum(r["flops"] for r in eng.profile)
layers = {}
for r in eng.profile:
    key = (tuple(r["shape"]), r["passes"], "simt" if r["impl"] == _lib.IMPL_SIMT else "tc")
    cur = layers.setdefault(key, [0, 0.0, 0.0])
    cur[0] += 1; cur[1] += r["ev"][0].elapsed_time(r["ev"][1]); cur[2] += r["flops"]
nconv = len(eng.profile)
eng.profile = None
model.use_graphs = True
for _ in range(3):
    xhat = model.decode(model.encode(x))
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
enc, dec = [], []
for _ in range(args.steps):
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    codes = model.encode(x)
    e1.record()
    xhat = model.decode(codes)
    e2.record()
    torch.cuda.synchronize()
    enc.append(e0.elapsed_time(e1)); dec.append(e1.elapsed_time(e2))
enc.sort(); dec.sort()
e, d = enc[len(enc) // 2], dec[len(dec) // 2]
print(json.dumps({"model": f"Compressor({args.channel}, {args.m}, {k})", "n": args.n, "hw": args.hw, "encode_ms": e,
                  "decode_ms": d, "mpix_s": args.n * args.hw * args.hw / (e + d) / 1e3, "conv_tflop_per_step": flops / 1e12,
                  "alg_tflops": flops / (e + d) / 1e9, "conv_launches": nconv, "finite": bool(torch.isfinite(xhat).all()),
                  "device_error_flag": int(_lib.load().mcq_device_error_flag()),
                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))
if args.layers:
    tot = sum(v[1] for v in layers.values())
    print(f"# eager conv time {tot:.1f} ms; (n,h,w,cin,cout,k,stride) passes impl: count, ms, share, executed TFLOP/s")
    for key, (cnt, ms, fl) in sorted(layers.items(), key=lambda kv: -kv[1][1])[:args.layers]:
        print(f"# {key[0]} p{key[1]} {key[2]}: {cnt:4d} {ms:8.2f} ms {ms / tot:6.1%} {fl * key[1] / ms / 1e9:8.1f}")
