"""Measurement script (not a test): the reference algorithm as eager PyTorch on the SAME GPU -- i.e. what the reference's
own code path (op-by-op ATen over cuDNN/cuBLAS) achieves on a B200 -- next to this framework.
The reference package itself cannot be imported on the GPU box, so its functional restatement (oracle/mcquic_oracle.py,
pinned bit-identical to the reference on CPU) is run on the CUDA device with cuDNN, with TF32 on (PyTorch's default for
convolutions) and off (true fp32, the only setting whose code indices are comparable).

    python tests/measure_eager_gpu.py [batch]
    python tests/measure_eager_gpu.py neon [batch] [hw]     # the a800_16 Neon tokenizer (tools/bench_neon.py's model)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform  # noqa: E402
from oracle import mcquic_oracle as O  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "neon":
    from mcquic_b200 import Neon
    from mcquic_b200.utils.synthetic import synthetic_block_state
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    hw = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    size = [16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1]
    template = Neon(256, 4096, size, False).state_dict()
    sd = {k: v.cuda() for k, v in synthetic_block_state(template, "neon.bench", seed=0).items()}
    x = uniform((batch, 3, hw, hw), "neon.bench.image", 0).cuda()
    out = {}
    with torch.no_grad():
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = True
            for _ in range(2):
                codes = O.neon_encode(sd, x, size)
                xh = O.neon_decode(sd, codes, size)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 3
            e[0].record()
            for _ in range(reps):
                codes = O.neon_encode(sd, x, size)
            e[1].record()
            for _ in range(reps):
                xh = O.neon_decode(sd, codes, size)
            e[2].record()
            torch.cuda.synchronize()
            te, td = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
            out["tf32" if tf32 else "fp32"] = {"encode_ms": te, "decode_ms": td, "images_per_s": batch / (te + td) * 1e3}
    print(json.dumps({"eager_torch_gpu_neon_a800_16": out, "batch": batch, "hw": hw, "gpu": torch.cuda.get_device_name(0)}))
    sys.exit(0)

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = [8192, 2048, 512]
sd = {k: v.cuda() for k, v in synthetic_state_dict(128, 1, K, seed=0).items()}
x = uniform((batch, 3, 256, 256), "bench.image.0", 0).cuda()
out = {}
with torch.no_grad():
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        for _ in range(3):
            codes = O.encode(sd, x)
            xh = O.decode(sd, codes)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        e[0].record()
        for _ in range(reps):
            codes = O.encode(sd, x)
        e[1].record()
        for _ in range(reps):
            xh = O.decode(sd, codes)
        e[2].record()
        torch.cuda.synchronize()
        te, td = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
        out["tf32" if tf32 else "fp32"] = {"encode_ms": te, "decode_ms": td,
                                           "mpix_s": batch * 256 * 256 / (te + td) / 1e3}
print(json.dumps({"eager_torch_gpu": out, "batch": batch, "gpu": torch.cuda.get_device_name(0)}))
