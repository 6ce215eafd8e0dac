#!/usr/bin/env python
"""Headline benchmark: encode+decode MPix/s at qp=1, batch 64x3x256x256 per GPU (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch: `model.encode(x)` (analysis transform + 3-level VQ, code
histogram fused into the VQ launch) then `model.decode(codes)` (gather + synthesis transform), plus -- at N>1 --
the single NCCL all-gather of the int32 code histogram (the path's only exchange step).  Each rank processes its
own batch of 64 images (weak scaling; images are independent).  Prints ONE JSON line on rank 0.

value      device time of K steps (CUDA events per step on the launching stream, L2 flushed between steps,
           max over ranks), inputs resident in HBM.
e2e        same steps through the public API with pinned HOST buffers: encode(host batch) (chunked H2D overlapping the
           first layers), D2H of the codes, decode(codes, out=host batch) (chunked D2H overlapping the last layers) --
           all inside the timed region.
roofline   every tcgen05 convolution launch of one step bracketed by CUDA events (eager pass after the timed
           region): achieved = sum(algorithmic FLOPs) / sum(durations) against the measured dense bf16/fp16 peak.
cpu_baseline / --impl reference: the oracle restatement of the reference's PyTorch CPU path (oracle/mcquic_oracle.py;
           /root/reference itself cannot travel to the GPU box) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNEL, M, K = 128, 1, [8192, 2048, 512]   # qp=1
BATCH, H, W = 64, 256, 256
STRONG_TOTAL = 512                           # BASELINE configs[3]: qp=1 batch 512 sharded over the ranks
ALG_GFLOP_PER_IMAGE = 89.44                  # encode 37.33 + decode 52.11 at 256x256
METRIC = "encode+decode MPix/s at qp=1, batch 64x3x256x256"
# one workload string for both arms (the driver pairs the `--impl reference` line with ours by metric and config)
WORKLOAD = "qp=1 Compressor(128,1,[8192,2048,512]) encode+decode, batch 64x3x256x256 per GPU"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            return json.load(fp), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


CPU_SAMPLE = 8      # images per CPU step: the SAME bounded sample for `cpu_baseline` and for `--impl reference`


def _physical_cores() -> int:
    """physical cores this process may run on (k = physical cores of the box, stated): distinct
    (package, core) pairs of /proc/cpuinfo restricted to the affinity mask; falls back to psutil / os.cpu_count()"""
    try:
        allowed = os.sched_getaffinity(0)
        cores, cpu, phys = set(), None, 0
        with open("/proc/cpuinfo") as fp:
            for line in fp:
                key, _, val = line.partition(":")
                key, val = key.strip(), val.strip()
                if key == "processor":
                    cpu = int(val)
                elif key == "physical id":
                    phys = int(val)
                elif key == "core id" and cpu in allowed:
                    cores.add((phys, int(val)))
        if cores:
            return len(cores)
    except Exception:
        pass
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def _cpu_oracle_throughput(sample_images: int, repeats: int):
    """encode+decode MPix/s of the CPU oracle on `sample_images` synthetic 256x256 images, best of `repeats`."""
    import torch
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    from oracle import mcquic_oracle as oracle
    cores = _physical_cores()
    torch.set_num_threads(cores)
    sd = synthetic_state_dict(CHANNEL, M, K, seed=0)
    x = uniform((sample_images, 3, H, W), "bench.image", 0)
    oracle.decode(sd, oracle.encode(sd, x[:1]))  # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        codes = oracle.encode(sd, x)
        oracle.decode(sd, codes)
        best = min(best, time.perf_counter() - t0)
    return sample_images * H * W / best / 1e6, best, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU PyTorch path (oracle restatement) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = CPU_SAMPLE
    times = []
    import torch
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    from oracle import mcquic_oracle as oracle
    cores = _physical_cores()
    torch.set_num_threads(cores)
    sd = synthetic_state_dict(CHANNEL, M, K, seed=0)
    x = uniform((sample, 3, H, W), "bench.image", 0)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.decode(sd, oracle.encode(sd, x))
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = sample * H * W * len(times) / total / 1e6
    desc = (f"{sample} of the 64 images of a step per CPU step (CPU time scales linearly in batch, SURVEY.md section 8d), "
            f"{cores} threads = physical cores of {os.cpu_count()} logical CPUs")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu": 64, "arm": "the reference's CPU PyTorch fp32 path on the host cores",
                   "sample": desc},
        "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": torch.get_num_threads(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def _by_class(rows, peak):
    """rows: [{"shape": (n,h,w,cin,cout,k,stride), "passes", "flops", "us"}] of the tcgen05 conv launches of one step ->
    per class of layer: launches, ms, executed TFLOP/s (flops x passes / time) and its fraction of the measured peak."""
    classes = {}
    for r in rows:
        n, h, w, cin, cout, k, stride = r["shape"]
        big = (h // stride) * (w // stride) >= 1024
        name = ("1x1 (GDN / IGDN / gate)" if k == 1 else
                ("3x3 on >= 32x32 maps, %d-pass" % r["passes"]) if big else "3x3 on <= 16x16 maps (latency-bound)")
        c = classes.setdefault(name, {"launches": 0, "ms": 0.0, "executed_tflop": 0.0})
        c["launches"] += 1
        c["ms"] += r["us"] * 1e-3
        c["executed_tflop"] += r["flops"] * r["passes"] / 1e12
    for c in classes.values():
        c["executed_tflops"] = c.pop("executed_tflop") / (c["ms"] * 1e-3) if c["ms"] > 0 else 0.0
        c["frac_of_peak"] = c["executed_tflops"] / peak
    return classes


def _dominant_in_graph(dev, peak):
    """The dominant kernel (conv_pair_kernel<3>, 64 x 64 x 64 x 128 -> 128) as it runs inside the timed step: a chain of
    dependent launches captured in ONE CUDA graph, CUDA events around the replay, L2 flushed before it; per-launch time =
    replay time / launches.  Two epilogue kinds as in the model: plane -> plane (first conv of a ResidualBlock) and
    fp32 residual in, fp32 + SiLU planes out (second conv).  The per-launch events of the eager leg add the launch gap and
    an idle pipeline at every kernel boundary; this figure has neither."""
    import torch
    from mcquic_b200.engine import Act, Engine, pack_conv
    L, n, hw, c = 10, BATCH, 64, CHANNEL
    eng = Engine("tcgen05")
    eng.chain = False
    eng.passes = 3
    g = torch.Generator().manual_seed(0)
    packs = [pack_conv(((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).to(dev), torch.zeros(c, device=dev),
                       1, 0, dev) for _ in range(L)]
    x = (torch.randn(n, hw, hw, c, generator=g) * 0.5).to(dev)
    hi = x.to(torch.float16)
    a0 = (hi, ((x - hi.float()) * 2048.0).to(torch.float16))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for kind in ("plane_to_plane", "residual_fp32_out"):
        def body():
            a, act, res = a0, Act(n, hw, hw, c), x
            for i in range(L):
                if kind == "plane_to_plane":
                    o = eng.conv(packs[i], a, act, {"silu"})
                else:
                    o = eng.conv(packs[i], a, act, {"f32", "silu"}, res1=res, res1_scale=0.5)
                    res = o.f32
                a, act = o.silu, o
            return a
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            body()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = body()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        us = 1e3 * sorted(ts)[len(ts) // 2] / L
        alg = 2.0 * n * hw * hw * c * c * 9 / us / 1e6
        out[kind] = {"us_per_launch": us, "alg_tflops": alg, "frac": alg / peak, "executed_tflops": 3 * alg,
                     "executed_frac": 3 * alg / peak}
        del gr, keep
    out["what"] = (f"{L} dependent launches of the dominant shape (n={n}, {hw}x{hw}, {c}->{c}, 3-pass) in one CUDA graph, "
                   "median of 5 replays / launches, L2 flushed before each replay")
    del flush
    torch.cuda.empty_cache()
    return out


def _eager_gpu_baseline(dev, x_dev):
    """The reference's own GPU path on THIS GPU in THIS run: its algorithm executed op by op through ATen / cuDNN / cuBLAS
    (the functional restatement oracle/mcquic_oracle.py on CUDA tensors -- the reference package cannot be imported on the
    GPU box; the restatement is bit-identical to it on CPU).  Two settings: PyTorch's default (TF32 convolutions; its code
    indices differ from the fp32 reference's) and true fp32 (`allow_tf32 = False`: the only setting whose codes are
    comparable, and the one this framework matches bit for bit).  A reported baseline, like cpu_baseline: nothing of the
    product runs through it."""
    import torch
    from mcquic_b200.utils.synthetic import synthetic_state_dict
    from oracle import mcquic_oracle as oracle
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(CHANNEL, M, K, seed=0).items()}
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        with torch.no_grad():
            for name, tf32 in (("tf32_default", True), ("fp32", False)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.benchmark = True
                for _ in range(2):
                    oracle.decode(sd, oracle.encode(sd, x_dev))
                torch.cuda.synchronize()
                reps = 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    oracle.decode(sd, oracle.encode(sd, x_dev))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                out[name] = {"ms_per_step": ms, "value": x_dev.shape[0] * H * W / ms / 1e3, "unit": "MPix/s"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    del sd
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mcquic_b200 import Compressor, _lib
    from mcquic_b200.dist import gather_histograms
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform

    # Libraries write banners to fd 1 (NCCL prints "NCCL version ..." there on communicator creation): everything
    # before the result goes to stderr, so that stdout carries exactly ONE JSON line.
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    _lib.apply_options(args.opt)      # A/B knobs (mcq_set_option), e.g. --opt direct_epi=1; empty = the library defaults
    model = Compressor(CHANNEL, M, K).eval()
    model.load_state_dict(synthetic_state_dict(CHANNEL, M, K, seed=0))
    model = model.to(dev)
    x_host = uniform((BATCH, 3, H, W), f"bench.image.{rank}", 0).pin_memory()
    x_dev = x_host.to(dev)
    hist_total = sum(M * k for k in K)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        hist = torch.zeros(hist_total, dtype=torch.int32, device=dev)
        codes = model.encode(x_dev, hist=hist)
        ghist = gather_histograms(hist) if world > 1 else hist
        xhat = model.decode(codes)
        return codes, xhat, ghist

    # e2e leg: uint8 RGB images in pinned host memory, as the reference's user-facing flow holds them (demo.py:48,109-134:
    # read_image -> uint8 -> compressImage ... decompressImage -> DeTransform -> uint8): the same 64 images quantised to
    # 8 bits; the input transform / DeTransform run inside the first / last kernel
    x_host_u8 = ((x_host + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).pin_memory()
    xhat_host = torch.empty((BATCH, 3, H, W), dtype=torch.uint8).pin_memory()
    codes_host = [torch.empty((BATCH, M, H >> (4 + l), W >> (4 + l)), dtype=torch.int64).pin_memory() for l in range(len(K))]

    def step_e2e():
        # public API with HOST buffers: the pinned image batch goes in, codes and pixels come back to pinned host memory.
        # encode(host) / decode(out=host) stream the batch in chunks that overlap the first / last layers.
        hist = torch.zeros(hist_total, dtype=torch.int32, device=dev)
        codes = model.encode(x_host_u8, hist=hist)
        if world > 1:
            gather_histograms(hist)
        for dst, src in zip(codes_host, codes):
            dst.copy_(src, non_blocking=True)
        model.decode(codes, out=xhat_host)
        torch.cuda.current_stream().synchronize()
        return int(xhat_host[0, 0, 0, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """sum of per-step device durations (ms); L2 flushed (untimed) before every step"""
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    barrier()
    launches_before = _lib.launch_count() + model.graph_launches
    with ClockSampler(local_rank) as clocks:
        barrier()
        prof_range = bool(os.environ.get("MCQ_CUDA_PROFILER_RANGE"))   # `ncu --profile-from-start off` captures only this
        if prof_range:
            torch.cuda.cudart().cudaProfilerStart()
        ms = timed(step_device, args.steps)
        if prof_range:
            torch.cuda.cudart().cudaProfilerStop()
        barrier()
        launches = _lib.launch_count() + model.graph_launches - launches_before
        ms_e2e = timed(step_e2e, args.steps)
        barrier()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    pix = world * BATCH * H * W * args.steps
    value = pix / (ms * 1e-3) / 1e6
    e2e_value = pix / (ms_e2e * 1e-3) / 1e6

    out = None
    roofline = None
    if rank == 0 and not args.no_roofline:
        # ---- roofline leg: one eager step with every conv launch bracketed by events (single stream)
        peaks, peak_src = _peaks()
        model.use_graphs = False
        model.engine.multistream = False
        prof = []
        model.engine.profile = prof
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
        codes = model.encode(x_dev)
        model.decode(codes)
        ev_b.record()
        torch.cuda.synchronize()
        eager_ms = ev_a.elapsed_time(ev_b)     # the same single-stream eager step the per-launch events sit in
        model.engine.profile = None
        model.use_graphs = True
        model.engine.multistream = True
        others = [p for p in prof if "other" in p]          # stem, VQ, gathers, histogram, layout kernels
        prof = [p for p in prof if "other" not in p]        # convolution launches
        t_other = sum(p["ev"][0].elapsed_time(p["ev"][1]) for p in others) * 1e-3
        t_all = t_other + sum(p["ev"][0].elapsed_time(p["ev"][1]) for p in prof) * 1e-3
        if args.dump_profile:
            with open(args.dump_profile, "w") as fp:
                json.dump([{"shape": p["shape"], "passes": p["passes"], "flops": p["flops"], "epi": p.get("epi"),
                            "us": 1e3 * p["ev"][0].elapsed_time(p["ev"][1])} for p in prof] +
                          [{"other": p["other"], "us": 1e3 * p["ev"][0].elapsed_time(p["ev"][1])} for p in others], fp)
        tc = [p for p in prof if p["impl"] == _lib.IMPL_TCGEN05]
        rows = [{"shape": p["shape"], "passes": p["passes"], "flops": p["flops"],
                 "us": 1e3 * p["ev"][0].elapsed_time(p["ev"][1])} for p in tc]
        t_tc = sum(r["us"] for r in rows) * 1e-6
        f_tc = sum(r["flops"] for r in rows)
        f_exec = sum(r["flops"] * r["passes"] for r in rows)
        peak = peaks["bf16_tflops_sustained"]
        # the dominant kernel: conv_pair_kernel<3> = the 3-pass (fp32-grade) 3x3 stride-1 convolutions of the encoder on
        # >= 32x32 maps (CTA-pair tcgen05 kernel).  Algorithmic FLOPs per launch = 2 * n*h*w * cout * 9*cin (SURVEY 8d's
        # hook-counted figure restricted to these layers); duration = CUDA events around each launch on its stream.
        dom = [r for r in rows if r["passes"] == 3 and r["shape"][5] == 3 and r["shape"][6] == 1
               and r["shape"][1] * r["shape"][2] >= 1024]
        t_dom = sum(r["us"] for r in dom) * 1e-6
        f_dom = sum(r["flops"] for r in dom)
        traffic, traffic_src = None, None
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as fp:
                tj = json.load(fp)
            if "dominant_kernel_dram_bytes_per_launch" in tj:
                traffic = tj["dominant_kernel_dram_bytes_per_launch"]
                traffic_src = "static: " + tj.get("dominant_kernel_source", "profiles/traffic.json") + \
                              " (an ncu --set full capture of the same kernel and shape, NOT measured by this run)"
        achieved = f_dom / t_dom / 1e12 if t_dom > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": "conv_pair_kernel<3>: 3-pass split-fp16 3x3 stride-1 convolutions on >= 32x32 maps, "
                              "%d launches per step" % len(dom),
                    "launches": len(dom), "avg_launch_us": 1e6 * t_dom / max(1, len(dom)),
                    "alg_gflop_per_launch": f_dom / max(1, len(dom)) / 1e9,
                    "executed_tflops": 3 * achieved, "executed_frac": 3 * achieved / peak,
                    "kernel_ms_per_step": t_dom * 1e3, "kernel_share_of_step": t_dom / t_all,
                    "timing": "CUDA events around every launch of one single-stream eager step run right after the timed "
                              "region (inside the graph-replayed two-stream timed step individual launches cannot be "
                              "bracketed); shares are over the device time of ALL %d launches of that eager step (%.2f ms; "
                              "its wall time %.2f ms; the timed step %.2f ms)"
                              % (len(prof) + len(others), t_all * 1e3, eager_ms, ms / args.steps),
                    "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops_sustained (fp16 dense = bf16 dense; the "
                                   "kernel is timed inside a long step)",
                    "note": "achieved counts algorithmic (fp32-semantics) FLOPs; the 3-pass split-fp16 kernel executes 3x of "
                            "them on the tensor pipe, so frac is capped at 1/3 and executed_frac is the tensor-pipe figure",
                    "all_tcgen05_convs": {"launches": len(rows), "alg_tflops": f_tc / t_tc / 1e12,
                                          "alg_frac": f_tc / t_tc / 1e12 / peak, "executed_tflops": f_exec / t_tc / 1e12,
                                          "executed_frac": f_exec / t_tc / 1e12 / peak, "ms_per_eager_step": t_tc * 1e3,
                                          "share_of_eager_step": t_tc / t_all},
                    "whole_step": {"alg_tflops": BATCH * ALG_GFLOP_PER_IMAGE / 1e3 / (ms / args.steps * 1e-3),
                                   "alg_frac": BATCH * ALG_GFLOP_PER_IMAGE / 1e3 / (ms / args.steps * 1e-3) / peak,
                                   "basis": "5.724 algorithmic TFLOP per 64-image step / the timed ms_per_step"},
                    "by_layer_class": _by_class(rows, peak)}
        roofline["dominant_in_graph"] = _dominant_in_graph(dev, peak)
    # (the roofline leg above runs right after the timed region, before the heavier legs below heat the GPU / fill its memory:
    #  measured, the same launches take up to 35 % longer when timed after the batch-512 leg)
    # ---- strong-scaling leg (BASELINE configs[3]): 512 images in total, 512 / N per rank, same step (encode, the one
    # histogram all-gather, decode); device-resident inputs, L2 flushed, max over ranks.  At N = 1 this is a batch-512 step.
    strong_ms, strong_steps = None, 3
    per_rank = STRONG_TOTAL // world
    if STRONG_TOTAL % world == 0 and not args.no_strong:
        xs = uniform((per_rank, 3, H, W), f"strong.image.{rank}", 0).to(dev)

        def step_strong():
            hist = torch.zeros(hist_total, dtype=torch.int32, device=dev)
            codes = model.encode(xs, hist=hist)
            if world > 1:
                gather_histograms(hist)
            return model.decode(codes)

        for _ in range(2):
            step_strong()
        barrier()
        strong_ms = timed(step_strong, strong_steps)
        barrier()
        del xs
    t = torch.tensor([strong_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    strong_ms = float(t[0])
    if rank == 0:
        cpu = None
        gpu_base = None
        if world == 1:
            if not args.no_gpu_baseline:
                gpu_base = _eager_gpu_baseline(dev, x_dev)
                for leg in gpu_base.values():
                    leg["ours_over_it"] = value / leg["value"]
                gpu_base["what"] = ("the reference algorithm as eager PyTorch (ATen / cuDNN / cuBLAS) on this GPU in this "
                                    "run, batch 64x3x256x256, encode+decode, device-resident input: PyTorch's default "
                                    "(TF32 convolutions) and true fp32 (the setting whose code indices are comparable)")
            if not args.no_cpu_baseline:
                cpu_val, cpu_s, cores = _cpu_oracle_throughput(sample_images=CPU_SAMPLE, repeats=5)
                cpu = {"value": cpu_val, "unit": "MPix/s", "cores": cores, "kind": "port",
                       "sample": f"{CPU_SAMPLE} of the 64 images of a step, encode+decode, best of 5 ({cpu_s:.2f} s each), "
                                 f"oracle/mcquic_oracle.py (CPU PyTorch fp32), {cores} threads = physical cores "
                                 f"({os.cpu_count()} logical CPUs)"}
        out = {
            "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 split x3 (fp32-grade) encode / f16 x1 decode, f32 accumulate",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "images_per_gpu": BATCH, "l2": "256 MB flush before every timed step", "cuda_graphs": True,
                       "collective": "all_gather int32[10752] code histogram per step" if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": x_host_u8.numel(),
                    "d2h_bytes_per_step": xhat_host.numel() + sum(c.numel() * 8 for c in codes_host),
                    "ms_per_step": ms_e2e / args.steps,
                    "buffers": "pinned host uint8 RGB in (encode(uint8): input transform in the stem kernel), int64 codes "
                               "and uint8 RGB out (decode(out=uint8): DeTransform in the last kernel's epilogue), copies "
                               "inside the timed region"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "options": args.opt or None,
            "cpu_baseline": cpu,
            "gpu_baseline": gpu_base,
            "strong_scaling": None if not strong_ms else {
                "total_images": STRONG_TOTAL, "images_per_gpu": STRONG_TOTAL // world, "n_gpus": world,
                "ms_per_step": strong_ms / strong_steps, "value": STRONG_TOTAL * H * W / (strong_ms / strong_steps) / 1e3,
                "unit": "MPix/s", "steps": strong_steps,
                "what": "BASELINE configs[3]: qp=1, 512 images in total sharded over the ranks (512 / N each), encode + "
                        "histogram all-gather + decode, device-resident inputs, L2 flushed, max over ranks; the 1 -> 8 curve "
                        "is this value across the driver's N = 1, 2, 4, 8 runs"},
            "alg_tflops": pix * ALG_GFLOP_PER_IMAGE / (H * W) * 1e9 / (ms * 1e-3) / 1e12,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)          # the JSON line is the only thing this process ever writes to its real stdout
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dump-profile", default=None, help="write the per-conv-launch timing list of the roofline leg here")
    ap.add_argument("--no-strong", action="store_true", help="skip the 512-image strong-scaling leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU-oracle leg (A/B runs of tools/)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-launch roofline leg (A/B runs of tools/)")
    ap.add_argument("--opt", default="", help="library options name=value,... (mcq_set_option); default: none")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
