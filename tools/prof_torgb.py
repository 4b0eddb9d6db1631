"""Profiling driver: the 1x1 ToRGB convolutions with the fused tail (conv mode 2: bias, clamp, + upsampled previous image).
python tools/prof_torgb.py [cin cout res B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from invertavatar_b200 import runtime as rt
from invertavatar_b200 import stylegan2 as sg

dev = 'cuda'
cases = [(128, 96, 256, 8), (128, 32, 256, 8), (256, 96, 128, 8), (512, 96, 64, 8)]
if len(sys.argv) > 4:
    cases = [tuple(int(a) for a in sys.argv[1:5])]
for (cin, cout, res, B) in cases:
    torch.manual_seed(0)
    L = sg.ToRGBLayer(cin, cout, w_dim=512, conv_clamp=256).requires_grad_(False).to(dev)
    a = rt.new_split(B, res, res, L.pack().Cin_pad, dev)
    a.hi.copy_(torch.randn(a.hi.shape, device=dev).bfloat16())
    a.lo.copy_((torch.randn(a.lo.shape, device=dev) * 0.004).bfloat16())
    prev = torch.randn(B, res // 2, res // 2, cout, device=dev)
    for it in range(3):
        L.run_split(a, img_prev=prev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        L.run_split(a, img_prev=prev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = B * res * res * (cin * 4 + cout * 4) + B * (res // 2) ** 2 * cout * 4
    print(f'torgb {cin}->{cout} @{res} B{B}: {ms * 1000:.1f} us  {byts / ms / 1e6:.0f} GB/s algorithmic')
