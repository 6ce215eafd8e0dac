"""Multi-GPU use of the hot path: images are independent, so a batch shards over ranks with replicated weights
and NO data-path collective.  The only exchange step is the per-rank code histogram that feeds the entropy
model (the reference all-reduces one fp32 [m,k] count per level, mcquic/modules/entropyCoder.py:34-36):
here it is ONE all-gather of a single flat int32 buffer [sum_l m*k_l] (43 KB at qp=1), summed locally --
integer counts, so the result is independent of rank order.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` images; the first `total % world_size` ranks get one extra."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_histograms(local_hist: torch.Tensor, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """all-gather the flat int32 histogram of every rank and sum: [sum_l m*k_l] global counts."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_hist.clone()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local_hist) for _ in range(world)]
    dist.all_gather(parts, local_hist.contiguous(), group=group)
    return torch.stack(parts).sum(0, dtype=torch.int64).to(local_hist.dtype)


def sharded_encode(model, x_local: torch.Tensor, group: Optional["dist.ProcessGroup"] = None,
                   update_frequencies: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Encode this rank's shard; returns (local codes, GLOBAL flat histogram).  With `update_frequencies` the
    model's `_entropyCoder._freqEMA` gets the reference's EMA update (entropyCoder.py:38-43) from the global counts,
    identically on every rank."""
    q = model._quantizer
    total = q.hist_size()
    hist = torch.zeros(total, dtype=torch.int32, device=x_local.device)
    codes = model.encode(x_local, hist=hist)
    global_hist = gather_histograms(hist, group)
    if update_frequencies:
        q._entropyCoder.update(global_hist)
    return codes, global_hist
