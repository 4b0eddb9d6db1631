"""One batch-8 generator step (after one warm-up step) for ncu captures:  ncu ... python tools/prof_one.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth, triplane
from invertavatar_b200.triplane import TriPlaneGenerator
triplane.set_backbone_streams(False)
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
B = 8
z, cond, c, uv = synth.latents(B).cuda(), synth.frontal_camera(B).cuda(), synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        img = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=128, noise_mode='const', evaluation=True)['image']
    torch.cuda.synchronize()
print('done', float(img.abs().mean()))
