"""world_size-2 gloo test of the multi-GPU host logic (sharding + single histogram all-gather), on CPU with the
emulated C ABI: the global histogram must equal the single-process bincount, codes must equal the unsharded run."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emulator import EmulatedLib
    from mcquic_b200 import Compressor
    from mcquic_b200.dist import shard_bounds, sharded_encode
    from mcquic_b200.engine import Engine
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    cfg = (32, 2, [32, 16])
    model = Compressor(*cfg).eval()
    model.load_state_dict(synthetic_state_dict(*cfg, seed=0))
    model._engine = Engine(lib=EmulatedLib())
    x = uniform((5, 3, 128, 128), "dist.image", 0)      # 5 images over 2 ranks: ragged shards (3 + 2)
    lo, hi = shard_bounds(5, world, rank)
    codes, ghist = sharded_encode(model, x[lo:hi], update_frequencies=True)
    torch.save({"codes": codes, "hist": ghist, "freq": [f.clone() for f in model._quantizer._entropyCoder._freqEMA],
                "bounds": (lo, hi)}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_histogram_all_gather(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from mcquic_b200.dist import shard_bounds
    from mcquic_b200.utils.synthetic import synthetic_state_dict, uniform
    from oracle import mcquic_oracle as O
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert [p["bounds"] for p in parts] == [(0, 3), (3, 5)] == [shard_bounds(5, 2, r) for r in range(2)]
    sd = synthetic_state_dict(32, 2, [32, 16], seed=0)
    ref = O.encode(sd, uniform((5, 3, 128, 128), "dist.image", 0))
    for lv in range(2):
        assert torch.equal(torch.cat([p["codes"][lv] for p in parts]), ref[lv])
    exp = torch.cat([h.flatten() for h in O.code_histogram(ref, [32, 16])]).int()
    assert torch.equal(parts[0]["hist"], exp) and torch.equal(parts[1]["hist"], exp)
    # the reference's EMA update (entropyCoder.py:38-43) from the global counts, identical on both ranks
    off = 0
    for lv, k in enumerate([32, 16]):
        cnt = exp[off:off + 2 * k].reshape(2, k).float()
        off += 2 * k
        want = 0.1 * cnt / cnt.sum(-1, keepdim=True) + 0.9 * torch.ones(2, k) / k
        assert torch.allclose(parts[0]["freq"][lv], want) and torch.equal(parts[0]["freq"][lv], parts[1]["freq"][lv])


def _neon_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import neon_inputs
    from emulator import EmulatedLib
    from mcquic_b200 import Neon
    from mcquic_b200.dist import shard_bounds, sharded_encode
    from mcquic_b200.engine import Engine
    model, x = neon_inputs("neon_c64_plain", Neon)
    model._engine = Engine(lib=EmulatedLib())
    x = torch.cat([x, x.flip(-1), x.flip(-2)])               # 3 images over 2 ranks: shards of 2 + 1
    lo, hi = shard_bounds(3, world, rank)
    codes, ghist = sharded_encode(model, x[lo:hi], update_frequencies=True)
    torch.save({"codes": codes, "hist": ghist, "freq": [f.clone() for f in model._quantizer._entropyCoder._freqEMA]},
               os.path.join(out_dir, f"neon_rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_histogram_all_gather_neon(tmp_path):
    """the same exchange step for the Neon tokenizer: one codebook per level (m is a list upstream), histogram
    segment j = codes[j] (smallest level first), EMA 0.998 (VariousMCoder, entropyCoder.py:293-322)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    world, port = 2, _free_port()
    mp.spawn(_neon_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f"neon_rank{r}.pt") for r in range(world)]
    k, levels = 128, 2
    allc = [torch.cat([p["codes"][j] for p in parts]) for j in range(levels)]
    assert [tuple(c.shape) for c in allc] == [(3, 1, 8, 8), (3, 1, 8, 8)]
    exp = torch.cat([torch.bincount(c.flatten(), minlength=k) for c in allc]).int()
    assert torch.equal(parts[0]["hist"], exp) and torch.equal(parts[1]["hist"], exp)
    for j in range(levels):
        cnt = exp[j * k:(j + 1) * k].reshape(1, k).float()
        want = 0.002 * cnt / cnt.sum(-1, keepdim=True) + 0.998 * torch.ones(1, k) / k
        assert torch.allclose(parts[0]["freq"][j], want, atol=1e-7) and torch.equal(parts[0]["freq"][j], parts[1]["freq"][j])


def test_shard_bounds_cover_everything():
    from mcquic_b200.dist import shard_bounds
    for total in (0, 1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
