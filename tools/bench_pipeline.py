"""Batch-8 steps issued round-robin on n streams (frames are independent: two batches in flight fill each other's bubbles)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth
from invertavatar_b200.triplane import TriPlaneGenerator
B = int(os.environ.get('IA_B', '8'))
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
z, cond = synth.latents(B).cuda(), synth.frontal_camera(B).cuda()
cams, uvs = synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()


def step():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    return G.synthesis(ws, cams, {'uvcoords_image': uvs}, neural_rendering_resolution=128, noise_mode='const', evaluation=True)['image']


def run(nstreams, n=12, warm=4):
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(nstreams)] if nstreams > 1 else [cur]
    def go(k):
        outs = []
        for i in range(k):
            s = streams[i % len(streams)]
            if s is not cur: s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(step())
        for s in streams:
            if s is not cur: cur.wait_stream(s)
        return outs
    go(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = go(n)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, outs[-1]


out = {'batch': B}
with torch.no_grad():
    t1, ref = run(1)
    out['ms_1stream'] = t1
    for ns in (2, 3):
        t, img = run(ns)
        out[f'ms_{ns}streams'] = t
        out[f'maxdiff_{ns}'] = float((img - ref).abs().max())
print(json.dumps(out))
