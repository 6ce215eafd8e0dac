"""Write-only HBM bandwidth of this GPU (torch fill of 4 GiB, CUDA events, best of 10): the practical ceiling for a
kernel whose only HBM-sized stream is an output (the soft-logit VQ), next to MEASURED_PEAKS.json's copy figure."""
import json
import torch

buf = torch.empty(1 << 32, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(13):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    buf.zero_()
    e1.record()
    e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"write_only_GBps": buf.numel() / best / 1e6, "bytes": buf.numel(), "ms": best}))
