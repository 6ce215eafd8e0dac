"""Timing / ncu driver of the training-step convolution kernels at the headline layer shape (64 x 64 x 64, 128 -> 128, 3x3):
forward (1 pass), dgrad, wgrad (csrc/conv_wgrad.cuh).   python tools/prof_wgrad.py [--once] [n h w cin cout]
   ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -o gpurun_out/wgrad python tools/prof_wgrad.py --once"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from mcquic_b200 import autograd as A
from mcquic_b200.utils.synthetic import uniform

once = "--once" in sys.argv
nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
n, h, w, cin, cout = nums if len(nums) == 5 else (64, 64, 64, 128, 128)
conv = nn.Conv2d(cin, cout, 3, padding=1).cuda()
x = uniform((n, cin, h, w), "prof.x", 0).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = uniform((n, cout, h, w), "prof.g", 1).cuda().contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
reps = 1 if once else 10
tf, tb = 0.0, 0.0
for it in range(reps + (0 if once else 2)):
    flush.fill_(1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = A.conv2d(conv, x)
    e[1].record()
    y.backward(g)
    e[2].record()
    e[2].synchronize()
    if once or it >= 2:
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
fl = 2.0 * n * h * w * cin * cout * 9
print(json.dumps({"shape": [n, h, w, cin, cout], "forward_ms (split + conv)": tf / reps, "backward_ms (scale + split + dgrad + wgrad + bias)": tb / reps,
                  "forward_TFLOPs": fl / (tf / reps) / 1e9, "backward_TFLOPs": 2 * fl / (tb / reps) / 1e9}))
