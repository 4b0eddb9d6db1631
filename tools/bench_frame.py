"""Batch-1 per-frame driver (reenact_avatar_next3d.py:214 / eval_seq.py:212): eager vs whole-frame CUDA graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth
from invertavatar_b200.graphs import GraphedSynthesis
from invertavatar_b200.triplane import TriPlaneGenerator
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
z, cond = synth.latents(1).cuda(), synth.frontal_camera(1).cuda()
cams, uvs = synth.cameras(8).cuda(), synth.uvcoords_image(8).cuda()


def timed(fn, n=20, warm=3):
    for _ in range(warm): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
with torch.no_grad():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    out['synthesis_eager_ms'] = timed(lambda i: G.synthesis(ws, cams[i % 8:i % 8 + 1], {'uvcoords_image': uvs[i % 8:i % 8 + 1]}, noise_mode='const', evaluation=True))
    gs = GraphedSynthesis(G, ws, cams[:1], uvs[:1])
    out['synthesis_graph_ms'] = timed(lambda i: gs(cams[i % 8:i % 8 + 1], uvs[i % 8:i % 8 + 1]))
    tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    out['with_texture_eager_ms'] = timed(lambda i: G.synthesis_withTexture(ws, tex, cams[i % 8:i % 8 + 1], {'uvcoords_image': uvs[i % 8:i % 8 + 1]},
                                                                           static_feats=sta, noise_mode='const', evaluation=True))
    gt = GraphedSynthesis(G, ws, cams[:1], uvs[:1], texture_feats=tex, static_feats=sta)
    out['with_texture_graph_ms'] = timed(lambda i: gt(cams[i % 8:i % 8 + 1], uvs[i % 8:i % 8 + 1]))
out['config'] = 'batch 1, 512^2 frame, 128^2 x 48+48, 1xB200'
print(json.dumps(out))
