"""A tile-starved convolution on its own (the encoder UNets' regime: 256 -> 256 at 32^2 x 4 images, split-K) for ncu captures and
knob sweeps:  python tools/prof_small_conv.py [cin cout res B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import stylegan2 as sg

cin, cout, res, B = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 256, 32, 4)
torch.manual_seed(0)
L = sg.SynthesisLayer(cin, cout, w_dim=512, resolution=res, up=1).requires_grad_(False).to('cuda')
x = torch.randn(B, res, res, cin, device='cuda')
st = torch.randn(B, cin, device='cuda')
dc = torch.rand(B, cout, device='cuda')
hi, lo = rt.modsplit(x, st, C_pad=L.pack().Cin_pad)
a = rt.Split(hi, lo)
nxt = rt.new_split(B, res, res, max(64, cout), 'cuda')
for it in range(3):
    L.run_split(a, dc, noise_mode='const', e1=(nxt, dc))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    L.run_split(a, dc, noise_mode='const', e1=(nxt, dc))
e1.record(); torch.cuda.synchronize()
print(f'{cin}->{cout} @{res}^2 x{B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (back to back)')
