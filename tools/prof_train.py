"""Top CUDA kernels of one training step (torch.profiler / CUPTI).  python tools/prof_train.py [c32|a800_16] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from mcquic_b200 import Neon
from mcquic_b200.utils.synthetic import synthetic_block_state, uniform

which = sys.argv[1] if len(sys.argv) > 1 else "c32"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == "c32":
    model = Neon(32, 4096, [16, 8, 4, 2, 2], True)
else:
    model = Neon(256, 4096, [16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1, 1], False)
model.load_state_dict(synthetic_block_state(model.state_dict(), "train.bench", seed=0))
model = model.cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)
x = uniform((batch, 3, 512, 512), "train.bench.image.0", 0).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    xHat, _, _, _ = model(x)
    F.mse_loss(xHat, x).backward()
    opt.step()

for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
