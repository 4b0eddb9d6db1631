"""BASELINE configs[2] (eval_seq.py few-shot path): e4e encode (B=1) + AR_eval_forward (T=4, 128^2 x 48+48 renders inside) +
4 x synthesis_withTexture, on cuda:0, with the per-kernel device-time breakdown.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth, runtime as rt
from invertavatar_b200.encoder import inversionNet
from invertavatar_b200.triplane import TriPlaneGenerator

T = 4
dev = 'cuda'
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
torch.manual_seed(1)
net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).train().requires_grad_(False)
synth.randomize_encoder(net)
for u in (net.unet_encoder.triplane_unet, net.unet_encoder.texture_unet):
    u.input_layer.eval(); u.body.eval()
net = net.to(dev)
x, c, v = synth.encoder_inputs(T)
x = {k: t.to(dev) for k, t in x.items()}; c = c.to(dev); v = {k: t.to(dev) for k, t in v.items()}


def timed(fn, n=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    ms_enc, ws = timed(lambda: net.encode(x['image'][:1]))
    G_ = net.generator
    tex = G_.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    sta = G_.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    e4e = {'w': ws, 'texture': tex, 'static': sta}
    state = {'r': [None, None]}

    def ar():
        upd, r = net.AR_eval_forward(x, c, v, ws, state['r'], e4e_results=e4e, return_fake=False)
        state['r'] = r
        return upd
    ms_ar, upd = timed(ar)
    ms_frame, img = timed(lambda: G_.synthesis_withTexture(ws, upd['texture'], c[:1], {k: t[:1] for k, t in v.items()}, noise_mode='const',
                                                           static_feats=upd['static'], evaluation=True)['image'], n=8)
    rt.profile_begin()
    net.encode(x['image'][:1]); ar()
    rep = rt.profile_report()
tot = sum(r['ms'] for r in rep.values())
top = {k: round(r['ms'], 3) for k, r in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])[:10]}
print(json.dumps({'config': 'eval_seq few-shot path: encode B=1 + AR_eval_forward T=4 + per-frame synthesis_withTexture, 1xB200',
                  'encode_ms': ms_enc, 'ar_eval_forward_ms': ms_ar, 'synthesis_withTexture_ms_per_frame': ms_frame,
                  'frames_per_s_per_frame_driver': 1000.0 / ms_frame, 'device_ms_encode_plus_ar': tot, 'top_kernels_ms': top,
                  'algorithmic_gflop': {'encode': 119.2, 'ar_eval_forward_T4': 2674.7}}))
