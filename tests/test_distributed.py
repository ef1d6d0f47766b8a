"""CPU tier: the N>1 host logic (utterance sharding + the final gather) with world_size-2 gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from world_b200 import distributed as wd


def test_shard_helpers():
    for n in (0, 1, 7, 256, 2048):
        for w in (1, 2, 4, 8):
            got = [wd.shard_range(n, r, w) for r in range(w)]
            idx = [i for lo, hi in got for i in range(lo, hi)]
            assert idx == list(range(n))
    lens = [64000, 100, 32000, 32000, 500, 64000, 7]
    parts = wd.shard_by_length(lens, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(lens)))
    loads = [sum(lens[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 64000


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 5
        lo, hi = wd.shard_range(n, rank, world)
        # every "utterance" i is a ramp of length 10 + 3 i filled with i + 1
        rows = [torch.full((10 + 3 * i,), float(i + 1), dtype=torch.float64) for i in range(lo, hi)]
        smax = max(len(r) for r in rows)
        local = torch.zeros((len(rows), smax), dtype=torch.float64)
        for k, r in enumerate(rows):
            local[k, :len(r)] = r
        lens = torch.tensor([len(r) for r in rows], dtype=torch.int32)
        allr, alll = wd.gather_padded(local, lens)
        # evenly sharded, equal shapes: the payload is gathered straight into the result
        even = torch.full((2, 7), float(rank + 1), dtype=torch.float64)
        er, el = wd.gather_padded(even, torch.tensor([7, 5], dtype=torch.int32))
        assert er.shape == (2 * world, 7) and [int(v) for v in el] == [7, 5] * world
        assert all(bool((er[2 * r:2 * r + 2] == r + 1).all()) for r in range(world))
        q.put((rank, allr.numpy(), alll.numpy()))
    finally:
        dist.destroy_process_group()


def test_gather_padded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, lens in got:
        assert list(lens) == [10 + 3 * i for i in range(5)]
        for i in range(5):
            assert np.all(rows[i, :lens[i]] == i + 1) and np.all(rows[i, lens[i]:] == 0)
