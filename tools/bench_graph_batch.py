"""Batch-8 step (BASELINE configs[1]): eager G.synthesis vs one whole-step CUDA graph replay."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth
from invertavatar_b200.graphs import GraphedSynthesis
from invertavatar_b200.triplane import TriPlaneGenerator
B = int(os.environ.get('IA_B', '8'))
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
z, cond = synth.latents(B).cuda(), synth.frontal_camera(B).cuda()
cams, uvs = synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()


def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {'batch': B}
with torch.no_grad():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    out['eager_ms'] = timed(lambda: G.synthesis(ws, cams, {'uvcoords_image': uvs}, noise_mode='const', evaluation=True))
    gs = GraphedSynthesis(G, ws, cams, uvs)
    out['graph_ms'] = timed(lambda: gs(cams, uvs))
print(json.dumps(out))
