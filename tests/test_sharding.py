"""N>1 host logic on CPU: world-size-2 gloo processes shard a global frame batch and gather the results in order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from invertavatar_b200.parallel import gather_frames, render_sharded, shard_range


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, cnt in spans:
                assert first == pos
                pos += cnt
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        z = torch.randn(n_frames, 4, generator=g)
        cond = torch.randn(n_frames, 2, generator=g)
        c = torch.randn(n_frames, 3, generator=g)
        uv = torch.randn(n_frames, 5, generator=g)

        def fake_render(z, cond, c, uv):   # frame-wise function: result of frame i depends on frame i only
            return (z.sum(1) + cond.sum(1) * 2 + c.sum(1) * 3 + uv.sum(1) * 5).view(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()
        out = render_sharded(fake_render, z, cond, c, uv)
        want = fake_render(z, cond, c, uv)
        ok = tuple(out.shape) == tuple(want.shape) and torch.equal(out, want)
        same = gather_frames(torch.full((2, 3, 2, 2), float(rank)))
        ok = ok and same.shape[0] == 2 * world and all(float(same[2 * r, 0, 0, 0]) == r for r in range(world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_gloo_shard_and_gather():
    ctx = mp.get_context('spawn')
    for n_frames in (8, 7):   # even and ragged split
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert sorted(res) == [(0, True), (1, True)], res
