"""GPU tier, two devices: the one collective on the path (main.py:198-214 followed by the gather
north_star names) -- every rank decodes its shard of the utterances and gather_padded() assembles the audio on
every rank over NCCL.  Skipped on a single-GPU box."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, os.path.join(root, "python-world_b200"))
    import torch.distributed as dist
    from world_b200 import distributed as wd, main, synth_input
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        W = main.World(device=rank)
        n = 6
        lens = [16000, 12000, 9000, 16000, 7000, 14000]
        xs = np.zeros((n, 16000))
        for i in range(n):
            xs[i, :lens[i]] = synth_input.utterance(16000, 1.0, 2, i)[:lens[i]]
        lo, hi = wd.shard_range(n, rank, world)
        dat = W.encode_batch(16000, xs[lo:hi], n_samples=lens[lo:hi], f0_method="dio", device_resident=True)
        E = W.engine
        tp, nf = dat["temporal_positions"], dat["n_frames"]
        ylen = 16001
        y, out_len, _ = E.decode(tp, dat["f0"], dat["vuv"], dat["spectrogram"], dat["aperiodicity"], nf, 16000, ylen, seed=9)
        rows, ln = wd.gather_padded(y, out_len)
        # reference for the check: rank 0 decodes everything itself (same seed, same per-utterance generator streams
        # only for its own shard -- so compare shard by shard against what each rank produced)
        q.put((rank, lo, hi, y.cpu().numpy(), out_len.cpu().numpy(), rows.cpu().numpy(), ln.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_decode_then_gather_nccl_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full0, len0 = got[0][5], got[0][6]
    assert np.array_equal(full0, got[1][5]) and np.array_equal(len0, got[1][6])  # identical on every rank
    assert len(len0) == 6
    for rank, lo, hi, y, ol, _, _ in got:  # rows lo..hi of the gathered array are that rank's audio
        for k in range(hi - lo):
            assert len0[lo + k] == ol[k] > 0
            assert np.array_equal(full0[lo + k, :ol[k]], y[k, :ol[k]])
            assert np.all(np.isfinite(y[k, :ol[k]])) and np.max(np.abs(y[k, :ol[k]])) > 1e-3
