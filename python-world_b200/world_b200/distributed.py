"""Multi-GPU plumbing: utterances are independent, so a batch is sharded by utterance across ranks (one
process per GPU, torch.distributed) with NO collective on the data path.  The only collective is an optional
final gather of results (decoded audio, or the small F0 arrays) onto every rank / rank 0 -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  SURVEY.md section 8e.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block of ceil(n/world) utterances per rank (the last ranks may get fewer or none)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def shard_by_length(lengths, world_size):
    """Length-balanced assignment: utterances sorted by length (longest first) are dealt to the currently
    lightest rank, so every rank holds a similar number of samples and similar Harvest buffer sizes.
    Returns a list of index lists, one per rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(v) for v in out]


def _all_gather(out, inp, group=None):
    """out [world * n, ...] <- every rank's inp [n, ...].  NCCL gathers straight into the flat result; other backends
    (gloo in the CPU tests) go through views of it."""
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, inp, group=group)
    else:
        world = dist.get_world_size(group)
        dist.all_gather(list(out.view((world,) + tuple(inp.shape)).unbind(0)), inp, group=group)


def gather_padded(local, local_len, group=None):
    """All-gather of per-rank padded rows.  local [b_r, S_r] (any float dtype), local_len [b_r] int32.
    Ranks may hold different b_r and S_r; returns (rows [sum b_r, max S], lengths [sum b_r]) in rank order,
    identical on every rank.  One all_gather of the shapes, then one all_gather_into_tensor of the payload
    straight into the result (no staging copies when every rank holds the same shape, the usual case of an
    evenly sharded batch) and one of the lengths."""
    world = dist.get_world_size(group)
    dev = local.device
    shape = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=dev)
    shapes = torch.empty((world, 2), dtype=torch.int64, device=dev)
    _all_gather(shapes.view(-1), shape, group)
    shapes = shapes.cpu()
    bs = [int(v) for v in shapes[:, 0]]
    bmax, smax = max(bs), int(shapes[:, 1].max())
    lens_local = local_len.to(torch.int32)
    if all(b == bmax for b in bs) and int(shapes[:, 1].min()) == smax:
        rows = torch.empty((world * bmax, smax), dtype=local.dtype, device=dev)
        lens = torch.empty(world * bmax, dtype=torch.int32, device=dev)
        _all_gather(rows, local.contiguous(), group)
        _all_gather(lens, lens_local.contiguous(), group)
        return rows, lens
    pad = torch.zeros((bmax, smax), dtype=local.dtype, device=dev)
    pad[:local.shape[0], :local.shape[1]] = local
    plen = torch.zeros(bmax, dtype=torch.int32, device=dev)
    plen[:local.shape[0]] = lens_local
    rows = torch.empty((world, bmax, smax), dtype=local.dtype, device=dev)
    lens = torch.empty((world, bmax), dtype=torch.int32, device=dev)
    _all_gather(rows.view(world * bmax, smax), pad, group)
    _all_gather(lens.view(-1), plen, group)
    if all(b == bmax for b in bs):
        return rows.view(world * bmax, smax), lens.view(world * bmax)
    return torch.cat([rows[r, :bs[r]] for r in range(world)], dim=0), torch.cat([lens[r, :bs[r]] for r in range(world)], dim=0)
