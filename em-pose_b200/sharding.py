"""Window sharding across ranks (one process per GPU).

Inference needs no data-path collective: in eval mode nothing couples two windows (BatchNorm uses running statistics,
the shape mean is intra-window -- reference ``models.py:529-535`` -- and the LSTM state is per window), so a batch is
cut into contiguous shards, every rank runs its shard, and results are concatenated in rank order.

Training is data parallel over windows: every rank computes the gradients of its shard into ONE flat vector and a
single all-reduce (sum, then divide by the world size) gives every rank the same averaged gradient -- DDP semantics.
Without BatchNorm that equals the whole-batch gradient for equal shard sizes (all losses are batch means,
``models.py:646-674``); with BatchNorm in train mode the batch statistics are per shard (no SyncBN), as with
``torch.nn.parallel.DistributedDataParallel`` around the reference model.
"""


def allreduce_mean_(flat, dist_module, group=None):
    """In-place average of a flat gradient vector over all ranks: exactly one collective."""
    dist_module.all_reduce(flat, op=dist_module.ReduceOp.SUM, group=group)
    flat.div_(dist_module.get_world_size(group))
    return flat


def shard_range(n_windows, rank, world_size):
    """Contiguous [begin, end) of the windows owned by ``rank``; sizes differ by at most one."""
    base, extra = divmod(n_windows, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every per-window tensor of a dict (first dimension = windows); ``None`` entries pass through."""
    n = next(t.shape[0] for t in tensors.values() if t is not None)
    b, e = shard_range(n, rank, world_size)
    return {k: (None if t is None else t[b:e]) for k, t in tensors.items()}


def gather_windows(local, n_windows, dist_module, group=None):
    """All-gather per-window results of unequal shard sizes and return the full [n_windows, ...] tensor (any backend)."""
    import torch
    world = dist_module.get_world_size(group)
    sizes = [shard_range(n_windows, r, world) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist_module.all_gather(parts, pad, group=group)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], dim=0)
