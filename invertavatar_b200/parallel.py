"""Frame sharding across the GPUs of one box.  Frames are independent, weights are replicated, and the only
collective of the path is the gather of the final images; one process per GPU under torchrun."""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous, balanced split of ``n_frames`` over ``world`` ranks -> (first, count)."""
    base, extra = divmod(int(n_frames), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_frames(local, counts=None, group=None):
    """All-gather of per-rank image batches [n_r,3,H,W] into [sum n_r,3,H,W] in rank order.  Equal counts use one
    all_gather_into_tensor (NCCL on GPU); ragged counts pad to the maximum."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    if counts is None:
        counts = [local.shape[0]] * world
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    m = max(counts)
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, counts)], dim=0)


def render_sharded(render_fn, z, cond, c, uv, group=None):
    """Shard the global batch over the ranks, run ``render_fn(z, cond, c, uv) -> images`` on the local frames and gather.
    All inputs are the GLOBAL batch (identical on every rank)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = z.shape[0]
    counts = [shard_range(n, r, world)[1] for r in range(world)]
    first, cnt = shard_range(n, rank, world)
    sl = slice(first, first + cnt)
    img = render_fn(z[sl], cond[sl], c[sl], uv[sl])
    return gather_frames(img, counts, group)
