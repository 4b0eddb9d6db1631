"""Per-shape device time of every convolution launch of one bench step (steady state, CUDA events around each launch).
    IA_PROF_DETAIL=1 python tools/prof_layers.py"""
import os, sys
os.environ.setdefault('IA_PROF_DETAIL', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth, runtime as rt
from invertavatar_b200.triplane import TriPlaneGenerator, set_backbone_streams
set_backbone_streams(False)      # one stream: per-launch event times must not overlap
B = 8
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
z, cond, c, uv = synth.latents(B).cuda(), synth.frontal_camera(B).cuda(), synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()
def step():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    return G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=128, noise_mode='const', evaluation=True)['image']
with torch.no_grad():
    for _ in range(3): step()
    torch.cuda.synchronize()
    rt.profile_begin()
    N = 3
    for _ in range(N): step()
    rep = rt.profile_report()
tot = sum(v['ms'] for v in rep.values()) / N
print(f'total {tot:.3f} ms/step')
import re
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms']):
    ms = v['ms'] / N
    if ms < 0.02: continue
    extra = ''
    m = re.match(r'ia_conv_tc\[t(\d+) (\d+)x(\d+) (\d+)->(\d+)\]', k)
    if m:
        t, gh, gw, ci, co = map(int, m.groups())
        fl = 2.0 * B * gh * gw * ci * co * t * v['launches'] / N
        extra = f'  {fl / (ms * 1e-3) / 1e12:7.1f} TF/s algorithmic (padded Cin)'
    print(f'{k:48s} {ms:8.3f} ms  x{v["launches"] / N:5.1f}{extra}')
