"""CPU precision probe for two candidate storage formats of the volume renderer, using the oracle only:
 (a) per-sample colours kept as fp16 between the decoder and the compositing (halves the per-ray scratch -> more ray-warps),
 (b) tri-planes stored as fp16 (halves the gather bytes).
Prints max-abs / PSNR of the final 512^2 image against the unmodified oracle on the same inputs (bar: 1e-3 / 50 dB)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from invertavatar_b200 import synth
from invertavatar_b200.triplane import TriPlaneGenerator
from oracle import triplane as o_tp, renderer as o_r

res, D = int(os.environ.get('RES', '64')), int(os.environ.get('D', '16'))
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(D, D)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
sd = {k: v.clone() for k, v in G.state_dict().items()}
z, cond, c, uv = synth.latents(1), synth.frontal_camera(1), synth.cameras(1), synth.uvcoords_image(1)
jit = synth.depth_jitter(1, res * res, D)


def run():
    with torch.no_grad():
        ws = o_tp.mapping(sd, z, cond, G.rendering_kwargs, truncation_psi=0.7, truncation_cutoff=14)
        return o_tp.synthesis(sd, ws, c, uv, G.rendering_kwargs, jit, evaluation=True, neural_rendering_resolution=res)['image']


def cmp(a, b):
    err = float((a - b).abs().max())
    mse = float(((a.double() - b.double()) ** 2).mean())
    return {'max_abs': err, 'psnr_db': float('inf') if mse == 0 else float(10 * np.log10(4.0 / mse))}


ref = run()
out = {'config': f'{res}^2 x {D}+{D}, batch 1, CPU oracle', 'image_range': [float(ref.min()), float(ref.max())]}
dec0 = o_r.osg_decoder
o_r.osg_decoder = lambda sdd, f: tuple((lambda rgb, sig: (rgb.half().float(), sig))(*dec0(sdd, f)))
out['colours_fp16'] = cmp(run(), ref)
o_r.osg_decoder = lambda sdd, f: tuple((lambda rgb, sig: (rgb.bfloat16().float(), sig))(*dec0(sdd, f)))
out['colours_bf16'] = cmp(run(), ref)
o_r.osg_decoder = dec0
sfp0 = o_r.sample_from_planes
o_r.sample_from_planes = lambda planes, coords, bw: sfp0(planes.half().float(), coords, bw)
out['planes_fp16'] = cmp(run(), ref)
o_r.sample_from_planes = sfp0
# (c) decoder MLP as single-pass fp16 tensor-core products (operands rounded to fp16, fp32 accumulation) instead of the 3-term split
import math
import torch.nn.functional as F
from oracle import stylegan2 as o_sg


def fc16(x, weight, bias):
    w = (weight * (1.0 / math.sqrt(weight.shape[1]))).half().float()
    return torch.addmm(bias.unsqueeze(0), x.half().float(), w.t())


def dec_single(sdd, feats, both=True):
    x = feats.mean(1)
    N, M, C = x.shape
    x = x.reshape(N * M, C)
    x = fc16(x, sdd['net.0.weight'], sdd['net.0.bias']) if both else o_sg.fully_connected(x, sdd['net.0.weight'], sdd['net.0.bias'])
    x = F.softplus(x)
    x = fc16(x, sdd['net.2.weight'], sdd['net.2.bias'])
    x = x.reshape(N, M, -1)
    return torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001, x[..., 0:1]


o_r.osg_decoder = dec_single
out['mlp_fp16_single_pass'] = cmp(run(), ref)
o_r.osg_decoder = lambda sdd, f: dec_single(sdd, f, both=False)
out['mlp_fp16_single_pass_layer2_only'] = cmp(run(), ref)
# everything together: fp16 planes + single-pass MLP + fp16 colours
o_r.sample_from_planes = lambda planes, coords, bw: sfp0(planes.half().float(), coords, bw)
o_r.osg_decoder = lambda sdd, f: tuple((lambda rgb, sig: (rgb.half().float(), sig))(*dec_single(sdd, f)))
out['all_three'] = cmp(run(), ref)
o_r.osg_decoder = dec0
o_r.sample_from_planes = sfp0
print(json.dumps(out))
