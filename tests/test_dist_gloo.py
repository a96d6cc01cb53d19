"""world_size-2 `gloo` test of the sharding / gather plumbing used by bench.py --gpus N (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cremage_b200.dist import full_batch_noise, gather_images, shard_batch, shard_range


def test_shard_range_partitions_exactly():
    for b in (0, 1, 7, 8, 9, 32):
        for w in (1, 2, 3, 8):
            spans = [shard_range(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gb = 5  # ragged on purpose: shards of 3 and 2
        x_T = full_batch_noise((gb, 4, 8, 8), seed=42, rank=rank, world=world)
        ctx = torch.arange(gb * 6, dtype=torch.float32).reshape(gb, 2, 3)
        (c_local,) = shard_batch([ctx], rank, world)
        assert x_T.shape[0] == c_local.shape[0]
        # "decode": a deterministic per-image function so the gather order is checkable
        img_local = (x_T.sum(dim=(1, 2, 3)) + c_local.sum(dim=(1, 2))).reshape(-1, 1, 1, 1).expand(-1, 2, 2, 3).contiguous()
        full = gather_images(img_local, gb)
        ref_x = torch.randn((gb, 4, 8, 8), generator=torch.Generator().manual_seed(42))
        ref = (ref_x.sum(dim=(1, 2, 3)) + ctx.sum(dim=(1, 2))).reshape(-1, 1, 1, 1).expand(-1, 2, 2, 3)
        ok = torch.equal(full, ref)
        # max-over-ranks timing reduction as bench.py does it
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ret[rank] = bool(ok and t.item() == world)
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
