#!/usr/bin/env python
"""BASELINE.json configs[4]: one TRAINING step (analysis + soft VQ + synthesis, forward + backward + optimizer) of the `Neon`
tokenizer, tokens/s, on 1 GPU or -- under torchrun -- N GPUs with DistributedDataParallel over NCCL.

    python tools/bench_train.py [--model c32|a800_16] [--batch 8] [--hw 512] [--steps 5] [--baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

model c32     = Neon(32, 4096, [16, 8, 4, 2, 2], denseNorm=True): SURVEY.md 8d's nearest constructible stand-in for
                configs/a800_8.yaml (which does not instantiate at reference HEAD); 1376 tokens per 512 x 512 image
model a800_16 = Neon(256, 4096, [16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1]) (configs/a800_16.yaml), 2384 tokens
step          = forward (BaseCompressor.forward, training mode) + MSE loss + backward + fused AdamW step; per-rank batch
                `--batch` (weak scaling, as the reference's DDP training), CUDA events, max over ranks
--baseline    = the same differentiable graph with every convolution through torch / cuDNN (TF32 allowed, as
                mcquic/train/utils.py sets it) instead of the tcgen05 kernels: the reference's GPU training path on this GPU
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="c32", choices=["c32", "a800_16"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--baseline", action="store_true")
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--opt", default="", help="library options, e.g. pair=1 (mcq_set_option)")
    args = ap.parse_args()
    from mcquic_b200 import Neon, _lib, autograd as A
    from mcquic_b200.utils.synthetic import synthetic_block_state, uniform

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.model == "c32":
        channel, k, size, dense = 32, 4096, [16, 8, 4, 2, 2], True
    else:
        channel, k, size, dense = 256, 4096, [16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1], False
    model = Neon(channel, k, size, dense)
    model.load_state_dict(synthetic_block_state(model.state_dict(), "train.bench", seed=0))
    model = model.to(dev).train()
    A.set_passes(args.passes)
    _lib.apply_options(args.opt)
    if args.baseline:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        A.conv2d = lambda conv, t: F.conv2d(t, conv.weight, conv.bias, conv.stride, conv.padding)
        A.conv2d_weights = lambda t, w, b, stride=1, owner=None: F.conv2d(t, w, b, stride, w.shape[-1] // 2)
    net = model
    use_graph = not args.no_graph
    if world > 1 and not use_graph:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True, capturable=use_graph)
    params = [p for p in model.parameters()]
    x = uniform((args.batch, 3, args.hw, args.hw), f"train.bench.image.{rank}", 0).to(dev)
    grid = args.hw // 8
    tokens, g = 0, grid
    last = size[0] * 2
    for s in size:                       # a level halves the grid when its size entry halves (quantizer.py:600-657)
        if s == last // 2:
            g //= 2
        tokens += g * g
        last = s

    def step():
        opt.zero_grad(set_to_none=use_graph is False)
        xHat, yHat, codes, logits = net(x)
        loss = F.mse_loss(xHat, x)
        loss.backward()
        if world > 1 and use_graph:
            # data-parallel gradient averaging inside the captured step: ONE all-reduce over the flattened gradients
            # (DistributedDataParallel's bucketed hooks do not replay from a CUDA graph; the arithmetic is the same)
            grads = [p.grad for p in params if p.grad is not None]
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat)
            flat.div_(world)
            for gr, new in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                gr.copy_(new)
        opt.step()
        return loss

    if use_graph:
        # the step is ~2 500 (c32) / ~6 600 (a800_16) kernel launches plus the autograd engine's Python: launched eagerly it
        # is host-bound (profiles/r2_prof_train_*.txt: 0.70 s of device time in a 0.99 s step), so it is captured once and
        # replayed as one CUDA graph.  Gradients keep their buffers between replays (zero_grad(set_to_none=False)).
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, args.warmup)):
                loss = step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = step()
        eager_step = step

        def step():                     # noqa: F811
            graph.replay()
            return static_loss
    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    before = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank == 0:
        print(json.dumps({
            "metric": "training step tokens/s (BASELINE configs[4])", "value": world * args.batch * tokens / ms * 1e3,
            "unit": "tokens/s", "n_gpus": world, "ms_per_step": ms, "images_per_s": world * args.batch / ms * 1e3,
            "tokens_per_image": tokens, "loss": float(loss), "impl": "torch/cuDNN convolutions (TF32)" if args.baseline
            else f"tcgen05 convolutions, {args.passes} pass(es)",
            "config": {"model": args.model, "channel": channel, "k": k, "size": size, "denseNorm": dense,
                       "batch_per_gpu": args.batch, "hw": args.hw, "loss": "MSE", "optimizer": "AdamW (fused)", "cuda_graph": use_graph,
                       "parallelism": f"ddp{world}" if world > 1 else "1 GPU"},
            "steps": args.steps, "warmup": args.warmup,
            "gpu_launches_per_step": (_lib.launch_count() - before) / args.steps,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    sys.stdout.flush()
    if world > 1:
        # a captured graph holds NCCL work: tearing the communicator down while it is alive hung the processes until
        # torchrun's timeout (measured); release the graph first, then leave without the collective teardown
        torch.cuda.synchronize()
        if use_graph:
            del graph
        os._exit(0)


if __name__ == "__main__":
    main()
