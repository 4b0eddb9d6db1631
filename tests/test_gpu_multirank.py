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
        ok = True
        consumed, wanted = [], []
        side = torch.cuda.Stream(device=dev)
        for step in range(5):
            # several steps, a consumer reading the gathered frames of step i on a SIDE stream while step i+1 is produced: the
            # double-buffered slots keep step i intact until the barrier after next (the consumer is joined before it)
            raw = torch.randn(B, H, W, Cc, generator=g).to(dev)
            bias = torch.randn(Cc, generator=g).to(dev)
            prev = torch.randn(B, H // 2, W // 2, Cc, generator=g).to(dev)
            torch.cuda.current_stream().wait_stream(side)          # consumer of step i-1 joined before barrier i closes
            with peer.sink():
                img = rt.torgb_finish(raw, bias, 256.0, prev, out_nchw=True)
            peer.barrier()
            want = torch.empty((world * B, Cc, H, W), device=dev)
            dist.all_gather_into_tensor(want, img.contiguous())
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                if rank == 1:
                    torch.cuda._sleep(20_000_000)                   # a slow consumer on one rank (~10 ms)
                consumed.append(peer.tensor.clone())
            wanted.append(want)
        torch.cuda.synchronize()
        ok = all(torch.equal(a, b) for a, b in zip(consumed, wanted))
        # a batch that does not match the slot size must raise instead of writing past the slot
        raised = False
        try:
            with peer.sink():
                rt.torgb_finish(torch.zeros(B + 1, H, W, Cc, device=dev), torch.zeros(Cc, device=dev), 256.0, None, out_nchw=True)
        except RuntimeError:
            raised = True
        q.put((rank, bool(ok and raised), bool(peer.mc_ptr)))
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
