"""Multi-GPU plumbing of the batched solver: one process per GPU, the batch index split contiguously over the
ranks.  The QPs are independent, so the data path has NO collective; `torch.distributed` is used
only for the optional final gather of the controls / statuses and for timing barriers.  Works with any backend
(`nccl` on the GPUs, `gloo` in the CPU tests)."""
import numpy as np


def shard_bounds(batch, world_size, rank):
    """[lo, hi) of the contiguous shard of `batch` instances owned by `rank`: sizes differ by at most one, the first
    `batch % world_size` ranks take the extra instance."""
    if world_size < 1 or not 0 <= rank < world_size or batch < 0:
        raise ValueError("need world_size >= 1, 0 <= rank < world_size, batch >= 0")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_arrays(arrays, world_size, rank):
    """This rank's rows of every [batch, ...] array (views, no copy)."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(a.shape[0], world_size, rank)
        out.append(a[lo:hi])
    return out


def gather_results(local, batch, group=None):
    """Optional final gather: every rank contributes its shard's rows of `local` ([shard, ...] numpy array) and gets
    the full [batch, ...] array back in instance order.  Ragged shards are padded to the largest shard for the
    all_gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    local = np.ascontiguousarray(local)
    if local.shape[0] != sizes[rank][1] - sizes[rank][0]:
        raise ValueError("local rows do not match this rank's shard")
    pad = np.zeros((biggest,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    use_cuda = dist.get_backend(group) == "nccl"
    if use_cuda:
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return np.concatenate([p.cpu().numpy()[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], axis=0)
