"""Profiling driver: a few representative modulated-convolution layers (used under ncu, or stand-alone for per-phase times:
IA_PROF_DETAIL=1 python tools/prof_conv.py 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import stylegan2 as sg

dev = 'cuda'
cases = [(128, 128, 512, 1, 8), (256, 128, 512, 2, 8), (512, 512, 64, 1, 8), (256, 256, 128, 1, 8), (512, 256, 128, 2, 8), (256, 128, 256, 2, 8)]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for (cin, cout, res, up, B) in cases:
    torch.manual_seed(0)
    L = sg.SynthesisLayer(cin, cout, w_dim=512, resolution=res, up=up).requires_grad_(False).to(dev)
    x = torch.randn(B, res // up, res // up, cin, device=dev)
    st = torch.randn(B, cin, device=dev)
    dc = torch.rand(B, cout, device=dev)
    hi, lo = rt.modsplit(x, st, C_pad=L.pack().Cin_pad)
    a = rt.Split(hi, lo)
    nxt = rt.new_split(B, res, res, max(64, cout), dev)
    for it in range(3):
        L.run_split(a, dc, noise_mode='const', e1=(nxt, dc))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        L.run_split(a, dc, noise_mode='const', e1=(nxt, dc))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2 * (res // up) ** 2 * cin * cout * 9 * B
    print(f'{cin}->{cout} @{res} up{up} B{B}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic')
    if os.environ.get('IA_PROF_DETAIL'):
        rt.profile_begin()
        for it in range(5):
            L.run_split(a, dc, noise_mode='const', e1=(nxt, dc))
        for k, v in sorted(rt.profile_report().items()):
            print(f'    {k:44s} {v["ms"] / 5 * 1000:8.1f} us')
