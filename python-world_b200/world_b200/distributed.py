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


def gather_padded(local, local_len, group=None):
    """All-gather of per-rank padded rows.  local [b_r, S_r] (any float dtype), local_len [b_r] int32.
    Ranks may hold different b_r and S_r; returns (rows [sum b_r, max S], lengths [sum b_r]) in rank order,
    identical on every rank.  One all_gather of the shapes and one of the padded payload."""
    world = dist.get_world_size(group)
    dev = local.device
    shape = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    bmax = int(max(int(s[0]) for s in shapes))
    smax = int(max(int(s[1]) for s in shapes))
    pad = torch.zeros((bmax, smax), dtype=local.dtype, device=dev)
    pad[:local.shape[0], :local.shape[1]] = local
    plen = torch.zeros(bmax, dtype=torch.int32, device=dev)
    plen[:local.shape[0]] = local_len.to(torch.int32)
    rows = [torch.zeros_like(pad) for _ in range(world)]
    lens = [torch.zeros_like(plen) for _ in range(world)]
    dist.all_gather(rows, pad, group=group)
    dist.all_gather(lens, plen, group=group)
    keep_r, keep_l = [], []
    for r in range(world):
        b = int(shapes[r][0])
        keep_r.append(rows[r][:b])
        keep_l.append(lens[r][:b])
    return torch.cat(keep_r, dim=0), torch.cat(keep_l, dim=0)
