"""BASELINE configs[4]: depth-sample sweep 16/48/96 x neural-resolution sweep 64/128/256 at batch 8 on one GPU.
Per point: frames/s of the whole step (CUDA events, after warm-up) and the device time of the fused volume renderer with its
algorithmic rates -- samples/s, MLP TFLOP/s (8320 FLOP per sample) and tri-plane gather TB/s (12 texels x 128 B per
sample, served by L1/L2: the planes are 25 MB per frame).  One JSON object on stdout."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth, runtime as rt
from invertavatar_b200 import triplane as tp
from invertavatar_b200.triplane import TriPlaneGenerator
B = 8
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
G = G.cuda()
z, cond, c, uv = synth.latents(B).cuda(), synth.frontal_camera(B).cuda(), synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()
points = []
with torch.no_grad():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    for N in (64, 128, 256):
        for D in (16, 48, 96):
            G.rendering_kwargs['depth_resolution'] = D
            G.rendering_kwargs['depth_resolution_importance'] = D
            step = lambda: G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=N, noise_mode='const', evaluation=True)['image']
            for _ in range(3): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            tp.set_backbone_streams(False)
            rt.profile_begin()
            for _ in range(3): step()
            rep = rt.profile_report()
            tp.set_backbone_streams(None)
            r_ms = rep['ia_render']['ms'] / 3
            samples = B * N * N * 2 * D
            points.append({'neural_res': N, 'depth_samples': [D, D], 'ms_per_step': round(ms, 3), 'frames_per_s': round(B / ms * 1e3, 1),
                           'render_ms': round(r_ms, 3), 'render_gsamples_per_s': round(samples / r_ms / 1e6, 2),
                           'render_mlp_tflops': round(samples * 8320 / r_ms / 1e9, 1),
                           'render_gather_tb_per_s': round(samples * 1536 / r_ms / 1e9, 2)})
print(json.dumps({'config': 'BASELINE configs[4]: batch 8, 512^2 output, 1xB200; synthesis only (mapping excluded)', 'points': points}))
