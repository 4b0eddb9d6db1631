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


class PeerFrameGather:
    """Gathered frame buffer [world*B,3,H,W] in symmetric (peer-mapped) memory: every rank's last ToRGB kernel writes its
    frames straight into slot ``rank`` of every rank's copy (``runtime.frame_sink``), so the image gather of SURVEY 8(e) is
    the epilogue of the producing kernel plus one cross-rank barrier, not a separate NCCL collective.  Uses the NVSwitch
    multicast mapping (multimem.st: one store, replicated in the switch) when the allocation has one, plain stores to the
    peer mappings otherwise.  Symmetric memory that cannot be set up raises; callers then fall back to ``gather_frames``
    (NCCL).

    The allocation is DOUBLE-BUFFERED: step i writes buffer i & 1 and ``barrier()`` flips.  Contract: after ``barrier()``
    returns, ``tensor`` is the gathered batch of the step just closed; it stays intact until this rank's stream reaches the
    NEXT ``barrier()`` -- a consumer (D2H copy, encoder) must be enqueued on the stream that calls ``barrier()``, or be waited
    for by it, before that next call.  Why that is enough: the buffer of step i is overwritten by the producing kernels of
    step i+2, which rank A launches after it passed barrier i+1 on the device, and a barrier completes only when every rank's
    stream has reached it -- i.e. after everything those ranks enqueued before it, their reads of step i included.  (With a
    single buffer, A's step i+1 kernel, ordered only after barrier i, could overwrite slot A in B's copy while B -- which
    enqueues its reads of step i AFTER barrier i -- is still reading.)"""

    def __init__(self, frames_per_rank, shape=(3, 512, 512), device=None, group=None, multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.per_rank = int(frames_per_rank) * int(shape[0]) * int(shape[1]) * int(shape[2])
        self.buffers = symm_mem.empty((2, self.world * frames_per_rank) + tuple(shape), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buffers, self.group)
        self.peer_ptrs = [int(q) for q in self.handle.buffer_ptrs]
        mc = 0
        try:
            if multicast and self.handle.has_multicast_support:
                mc = int(self.handle.multicast_ptr or 0)
        except Exception:
            mc = 0
        self.mc_ptr = mc
        self.buf_elems = self.world * self.per_rank
        self.cur = 0          # buffer the next sink() writes
        self.last = 0         # buffer closed by the most recent barrier()

    @property
    def tensor(self):
        """[world*B,3,H,W] gathered frames of the step closed by the most recent ``barrier()``."""
        return self.buffers[self.last]

    def sink(self):
        from . import runtime as rt
        off = self.cur * self.buf_elems + self.rank * self.per_rank
        return rt.frame_sink(self.peer_ptrs, off, self.mc_ptr, capacity=self.per_rank)

    def barrier(self):
        """All ranks' frame writes of this step are visible in every copy after this (enqueued on the current stream); the
        next ``sink()`` targets the other buffer."""
        self.handle.barrier(channel=0)
        self.last, self.cur = self.cur, self.cur ^ 1
