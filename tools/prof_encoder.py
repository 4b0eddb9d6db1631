"""Per-shape device time of the encoder path (encode + AR_eval_forward T=4): IA_PROF_DETAIL=1 python tools/prof_encoder.py"""
import os, sys, re
os.environ.setdefault('IA_PROF_DETAIL', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth, runtime as rt
from invertavatar_b200.encoder import inversionNet
from invertavatar_b200.triplane import TriPlaneGenerator, set_backbone_streams
set_backbone_streams(False)                                   # one stream: per-launch event times must not overlap
rt.side_streams = lambda device: (torch.cuda.current_stream(device),) * 2
T = 4
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
torch.manual_seed(1)
net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).train().requires_grad_(False)
for u in (net.unet_encoder.triplane_unet, net.unet_encoder.texture_unet):
    u.input_layer.eval(); u.body.eval()
net = net.cuda()
x, c, v = synth.encoder_inputs(T)
x = {k: t.cuda() for k, t in x.items()}; c = c.cuda(); v = {k: t.cuda() for k, t in v.items()}
with torch.no_grad():
    ws = net.encode(x['image'][:1])
    G_ = net.generator
    tex = G_.texture_backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    sta = G_.backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    e4e = {'w': ws, 'texture': tex, 'static': sta}
    r = [None, None]
    for _ in range(2):
        upd, r = net.AR_eval_forward(x, c, v, ws, r, e4e_results=e4e)
    torch.cuda.synchronize()
    which = sys.argv[1] if len(sys.argv) > 1 else 'ar'
    rt.profile_begin()
    if which == 'encode':
        net.encode(x['image'][:1])
    else:
        net.AR_eval_forward(x, c, v, ws, r, e4e_results=e4e)
    rep = rt.profile_report()
tot = sum(q['ms'] for q in rep.values())
print(f'{which}: total {tot:.3f} ms, {sum(q["launches"] for q in rep.values())} launches')
for k, q in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])[:60]:
    extra = ''
    m = re.match(r'ia_conv_tc\[t(\d+) (\d+)x(\d+) (\d+)->(\d+)\]', k)
    if m:
        t, gh, gw, ci, co = map(int, m.groups())
        extra = f'  {2.0 * gh * gw * ci * co * t * q["launches"] / (q["ms"] * 1e-3) / 1e12:7.1f} TF/s per image (x batch; padded Cin)'
    print(f'{k:48s} {q["ms"]:8.3f} ms  x{q["launches"]:4d}{extra}')
