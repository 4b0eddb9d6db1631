"""Two-rank NCCL check of the fused image gather (skipped on boxes with fewer than two GPUs; bench.py repeats the same
bit-for-bit comparison during its warm-up whenever it runs on more than one GPU)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from invertavatar_b200 import runtime as rt
    from invertavatar_b200.parallel import PeerFrameGather
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        B, Cc, H, W = 2, 3, 16, 16
        peer = PeerFrameGather(B, (Cc, H, W), device=dev)
        g = torch.Generator().manual_seed(100 + rank)
        raw = torch.randn(B, H, W, Cc, generator=g).to(dev)
        bias = torch.randn(Cc, generator=g).to(dev)
        prev = torch.randn(B, H // 2, W // 2, Cc, generator=g).to(dev)
        with peer.sink():
            img = rt.torgb_finish(raw, bias, 256.0, prev, out_nchw=True)
        peer.barrier()
        want = torch.empty((world * B, Cc, H, W), device=dev)
        dist.all_gather_into_tensor(want, img.contiguous())
        torch.cuda.synchronize()
        q.put((rank, bool(torch.equal(want, peer.tensor)), bool(peer.mc_ptr)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_fused_gather_matches_nccl_all_gather():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
