"""Profiling driver: the tensor-core attention of the one-shot encoder at its largest size (up4: 64^2 = 4096 tokens, 4 heads x 256).
python tools/prof_attention.py [tokens_side]     (used under ncu, or stand-alone for the device time)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from invertavatar_b200 import runtime as rt

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = 'cuda'
g = torch.Generator().manual_seed(0)
q = torch.randn(1, side, side, 1024, generator=g).to(dev)
kv = torch.randn(1, side, side, 2048, generator=g).to(dev)


def split(x):
    hi = x.bfloat16()
    return rt.Split(hi.contiguous(), (x - hi.float()).bfloat16().contiguous())


qs, kvs = split(q), split(kv)
for _ in range(3):
    rt.attention_tc(qs, kvs, 4, 256 ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    rt.attention_tc(qs, kvs, 4, 256 ** -0.5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2 * 2 * (side * side) ** 2 * 1024
print(f'attention_tc {side * side} tokens x 4 heads x 256: {ms * 1000:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s algorithmic, {3 * fl / ms / 1e9:.1f} TFLOP/s of issued MMAs')
