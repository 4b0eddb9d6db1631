"""CPU precision probe for the tensor-core operand format of every modulated convolution of the generator,
using the oracle only.  For each 3x3 layer (call order of one synthesis: texture backbone, static backbone, face backbone,
super-resolution) the final 512^2 image is rendered with THAT layer's operands rounded as a cheaper tensor-core scheme would
round them, everything else exact fp32, and compared with the unmodified oracle:

    f16x1   A = rn_f16(x*s), W = rn_f16(w)                      1 MMA per k-step   (single pass)
    f16a    A = x*s (hi+lo, ~22 bits), W = rn_f16(w)            2 MMAs             (hi*hi + lo*hi)
    f16w    A = rn_f16(x*s), W = w (hi+lo)                      2 MMAs             (hi*hi + hi*lo)
    bf16x3 / f16x3 are emulated exactly (3 convolutions) by --verify only.

Then a greedy mix is built under an error budget and verified on several frames.
    python tools/probe_conv_precision.py [--frames 3] [--budget 5e-4] > profiles/r2_conv_precision_probe.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from invertavatar_b200 import synth
from invertavatar_b200.triplane import TriPlaneGenerator
from oracle import ops as o_ops
from oracle import stylegan2 as o_sg
from oracle import triplane as o_tp

ap = argparse.ArgumentParser()
ap.add_argument('--res', type=int, default=128)
ap.add_argument('--depth', type=int, default=48)
ap.add_argument('--frames', type=int, default=3)
ap.add_argument('--budget', type=float, default=5e-4)
ap.add_argument('--modes', default='f16x1,f16a,f16w')
ap.add_argument('--mlp-fp16', action='store_true', help='also evaluate the renderer decoder MLP as single-pass fp16 products (IA_RENDER_MLP=fp16)')
ap.add_argument('--torgb-f16', action='store_true', help='eval-mix: the backbone ToRGB layers (1x1, no demodulation) single-pass fp16 as well')
ap.add_argument('--eval-mix', default='', help="skip the search: evaluate a preset ('backbones_f16x1') on --frames frames, with bf16x3 arithmetic elsewhere")
args = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)

torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(args.depth, args.depth)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
sd = {k: v.clone() for k, v in G.state_dict().items()}
kw = G.rendering_kwargs

PREC = {}          # conv call index -> mode
STATE = {'i': 0, 'layers': []}
_conv0 = o_ops.conv2d_resample


def rn(t, mode):
    return t.half().float() if mode.startswith('f16') else t.bfloat16().float()


def split2(t, mode):
    hi = rn(t, mode)
    return hi, rn(t - hi, mode)


def modconv(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True, flip_weight=True, fused_modconv=True):
    """The engine's formulation (fused_modconv=False, networks_stylegan2_new.py:70-79) with rounded tensor-core operands."""
    if not demodulate:       # ToRGB: 3-term (exact here) unless --torgb-f16; the super-resolution ToRGBs (3 output channels) are fp32 FMAs
        if STATE.get('torgb_f16') and weight.shape[0] > 3:
            B_, I_ = x.shape[0], weight.shape[1]
            a = x * styles.reshape(B_, I_, 1, 1)
            y = _conv0(rn(a, 'f16x1'), rn(weight, 'f16x1'), f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
            return y + noise if noise is not None else y
        return _mod0(x, weight, styles, noise=noise, up=up, padding=padding, resample_filter=resample_filter, demodulate=demodulate,
                     flip_weight=flip_weight, fused_modconv=fused_modconv)
    idx = STATE['i']
    STATE['i'] += 1
    B, (O, I, kh, kwd) = x.shape[0], weight.shape
    if len(STATE['layers']) <= idx:
        r_in = x.shape[-1]
        STATE['layers'].append({'idx': idx, 'cin': I, 'cout': O, 'res_in': r_in, 'up': up, 'gflop': 2 * r_in * r_in * I * O * 9 / 1e9})
    mode = PREC.get(idx)
    if mode is None:
        return _mod0(x, weight, styles, noise=noise, up=up, padding=padding, resample_filter=resample_filter, demodulate=demodulate,
                     flip_weight=flip_weight, fused_modconv=fused_modconv)
    w = weight.unsqueeze(0) * styles.reshape(B, 1, I, 1, 1)
    dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    a = x * styles.reshape(B, I, 1, 1)

    def conv(aa, ww):
        return _conv0(aa, ww, f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
    if mode in ('f16x1', 'bf16x1'):
        y = conv(rn(a, mode), rn(weight, mode))
    elif mode in ('f16a', 'bf16a'):
        y = conv(a, rn(weight, mode))
    elif mode in ('f16w', 'bf16w'):
        y = conv(rn(a, mode), weight)
    elif mode in ('f16x3', 'bf16x3'):
        ah, al = split2(a, mode)
        wh, wl = split2(weight, mode)
        y = conv(ah, wh) + conv(ah, wl) + conv(al, wh)
    else:
        raise ValueError(mode)
    y = y * dcoefs.reshape(B, O, 1, 1)
    return y + noise if noise is not None else y


_mod0 = o_sg.modulated_conv2d
o_sg.modulated_conv2d = modconv


def render(frame):
    STATE['i'] = 0
    z, cond, c, uv = synth.latents(1, frame), synth.frontal_camera(1), synth.cameras(1, frame), synth.uvcoords_image(1, frame)
    jit = synth.depth_jitter(1, args.res * args.res, args.depth, seed=7 + frame)
    with torch.no_grad():
        ws = o_tp.mapping(sd, z, cond, kw, truncation_psi=0.7, truncation_cutoff=14)
        return o_tp.synthesis(sd, ws, c, uv, kw, jit, evaluation=True, neural_rendering_resolution=args.res)['image']


def cmp(a, b):
    err = float((a - b).abs().max())
    mse = float(((a.double() - b.double()) ** 2).mean())
    return err, (float('inf') if mse == 0 else float(10 * np.log10(4.0 / mse)))


t0 = time.time()
refs = [render(f) for f in range(args.frames)]
layers = STATE['layers']
names = []
for net in ('texture', 'static', 'face'):
    names += [f'{net}.b4.conv1'] + [f'{net}.b{r}.conv{k}' for r in (8, 16, 32, 64, 128, 256) for k in (0, 1)]
names += ['sr.block0.conv0', 'sr.block0.conv1', 'sr.block1.conv0', 'sr.block1.conv1']
assert len(names) == len(layers), (len(names), len(layers))
for L, n in zip(layers, names):
    L['name'] = n
total_gflop = sum(L['gflop'] for L in layers)
print(f'{len(layers)} conv layers, {total_gflop:.1f} GFLOP/frame, reference render {time.time() - t0:.1f} s for {args.frames} frames', file=sys.stderr)

if args.eval_mix:
    assert args.eval_mix == 'backbones_f16x1'
    if args.mlp_fp16:
        import math
        import torch.nn.functional as F
        from oracle import renderer as o_r

        def fc16(x, weight, bias):
            w = (weight * (1.0 / math.sqrt(weight.shape[1]))).half().float()
            return torch.addmm(bias.unsqueeze(0), x.half().float(), w.t())

        def dec_single(sdd, feats):
            x = feats.mean(1)
            N, M, C = x.shape
            x = fc16(x.reshape(N * M, C), sdd['net.0.weight'], sdd['net.0.bias'])
            x = fc16(F.softplus(x), sdd['net.2.weight'], sdd['net.2.bias']).reshape(N, M, -1)
            return torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001, x[..., 0:1]
        refs = [render(f) for f in range(args.frames)]      # (references were rendered above with the exact decoder)
        o_r.osg_decoder = dec_single
    mix = {L['idx']: 'f16x1' for L in layers if not L['name'].startswith('sr.')}
    STATE['torgb_f16'] = bool(args.torgb_f16)
    full = {L['idx']: mix.get(L['idx'], 'bf16x3') for L in layers}
    res = []
    for f in range(args.frames):
        PREC.clear(); PREC.update(full)
        e, p = cmp(render(f), refs[f])
        res.append({'frame': f, 'max_abs': e, 'psnr_db': p})
        print(res[-1], file=sys.stderr)
    issued = sum(L['gflop'] * (1 if L['idx'] in mix else 3) for L in layers)
    print(json.dumps({'config': f'{args.res}^2 x {args.depth}+{args.depth}, batch 1, CPU oracle, {args.frames} frames', 'mix': 'every backbone 3x3 layer single-pass fp16 '
                      '(A = rn_f16(x*s), W = rn_f16(w)), super-resolution layers and ToRGB layers bf16x3', 'frames': res,
                      'total_gflop_per_frame': total_gflop, 'issued_gflop_mix': issued, 'issued_gflop_3term': 3 * total_gflop,
                      'algorithmic_over_issued': total_gflop / issued}, indent=1))
    sys.exit(0)

modes = args.modes.split(',')
cost = {'f16x1': 1, 'bf16x1': 1, 'f16a': 2, 'f16w': 2, 'bf16a': 2, 'bf16w': 2}
sens = {}
for L in layers:
    for m in modes:
        PREC.clear()
        PREC[L['idx']] = m
        e, p = cmp(render(0), refs[0])
        sens[(L['idx'], m)] = e
        L.setdefault('err', {})[m] = e
    print(L['name'], f"{L['gflop']:.1f} GF", {m: f"{L['err'][m]:.1e}" for m in modes}, file=sys.stderr)

# greedy: candidate moves (layer -> mode) ordered by error added per MMA-FLOP saved; accept while the MEASURED error of the
# mix (frame 0) stays under the budget
cands = []
for L in layers:
    for m in modes:
        saved = L['gflop'] * (3 - cost[m])
        cands.append((sens[(L['idx'], m)] / saved, L['idx'], m, saved))
cands.sort()
mix, mix_err = {}, 0.0
for _, idx, m, saved in cands:
    if idx in mix and cost[mix[idx]] <= cost[m]:
        continue
    trial = dict(mix)
    trial[idx] = m
    PREC.clear(); PREC.update(trial)
    e, p = cmp(render(0), refs[0])
    if e <= args.budget:
        mix, mix_err = trial, e
issued = sum(L['gflop'] * (cost[mix[L['idx']]] if L['idx'] in mix else 3) for L in layers)
PREC.clear(); PREC.update(mix)
verify = [cmp(render(f), refs[f]) for f in range(args.frames)]
# the same mix on top of the real bf16x3 arithmetic of the remaining layers (what the GPU computes)
full = {L['idx']: mix.get(L['idx'], 'bf16x3') for L in layers}
PREC.clear(); PREC.update(full)
verify_x3 = [cmp(render(f), refs[f]) for f in range(min(2, args.frames))]
PREC.clear(); PREC.update({L['idx']: 'bf16x3' for L in layers})
base_x3 = cmp(render(0), refs[0])
out = {'config': f'{args.res}^2 x {args.depth}+{args.depth}, batch 1, CPU oracle, {args.frames} frames', 'budget_max_abs': args.budget,
       'total_gflop_per_frame': total_gflop, 'issued_gflop_3term': 3 * total_gflop, 'issued_gflop_mix': issued,
       'issued_ratio': issued / (3 * total_gflop), 'algorithmic_over_issued': total_gflop / issued,
       'mix': {layers[i]['name']: m for i, m in sorted(mix.items())},
       'mix_err_frame0': mix_err, 'verify_frames': [{'max_abs': e, 'psnr_db': p} for e, p in verify],
       'verify_with_bf16x3_elsewhere': [{'max_abs': e, 'psnr_db': p} for e, p in verify_x3],
       'baseline_all_bf16x3': {'max_abs': base_x3[0], 'psnr_db': base_x3[1]},
       'layers': [{k: v for k, v in L.items()} for L in layers]}
print(json.dumps(out, indent=1))
