"""Which part of the conv epilogue costs what: one 3x3 layer shape timed with every combination of emitted outputs
(fp32 NHWC copy, operand of the next convolution, operand of the ToRGB layer) and with / without the noise input.
    python tools/prof_conv_variants.py [cin cout res B]..."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import stylegan2 as sg

dev = 'cuda'
args = [int(a) for a in sys.argv[1:]]
cases = [tuple(args[i:i + 4]) for i in range(0, len(args), 4)] or [(128, 128, 256, 8), (128, 128, 512, 8), (256, 256, 128, 8), (512, 512, 64, 8)]
for (cin, cout, res, B) in cases:
    torch.manual_seed(0)
    L = sg.SynthesisLayer(cin, cout, w_dim=512, resolution=res, up=1).requires_grad_(False).to(dev)
    x = torch.randn(B, res, res, cin, device=dev)
    st = torch.randn(B, cin, device=dev)
    dc = torch.rand(B, cout, device=dev)
    hi, lo = rt.modsplit(x, st, C_pad=L.pack().Cin_pad)
    a = rt.Split(hi, lo)
    nxt = rt.new_split(B, res, res, max(64, cout), dev)
    rgb = rt.new_split(B, res, res, max(64, cout), dev)
    fl = 2 * res ** 2 * cin * cout * 9 * B
    for want32, use1, use2, noise in [(0, 1, 0, 'none'), (0, 1, 0, 'const'), (0, 1, 1, 'const'), (1, 0, 1, 'const'), (1, 1, 1, 'const'), (1, 0, 0, 'const'), (1, 0, 0, 'none')]:
        kw = dict(noise_mode=noise, want32=bool(want32), e1=(nxt, dc) if use1 else None, e2=(rgb, dc) if use2 else None)
        for it in range(3):
            L.run_split(a, dc, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(5):
            L.run_split(a, dc, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{cin}->{cout} @{res} B{B} out32={want32} e1={use1} e2={use2} noise={noise}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TF/s')
