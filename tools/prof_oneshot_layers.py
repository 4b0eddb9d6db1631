"""Per-shape device time of every launch of one uvnet_new.inversionNet.forward (IA_PROF_DETAIL names)."""
import os, sys
os.environ.setdefault('IA_PROF_DETAIL', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import runtime as rt, synth
from invertavatar_b200.segformer import inversionNet
from invertavatar_b200.triplane import TriPlaneGenerator
dev = 'cuda'
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False)
synth.randomize_noise_and_wavg(G)
torch.manual_seed(1)
net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).eval().requires_grad_(False)
synth.randomize_encoder(net)
synth.randomize_by_name(net.unet_encoder)
net = net.to(dev)
x, c, v = synth.encoder_inputs(1)
x = {k: t.to(dev) for k, t in x.items()}; c = c.to(dev); v = {k: t.to(dev) for k, t in v.items()}
with torch.no_grad():
    ws = net.encode(x['image'])
    tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    e4e = {'w': ws, 'texture': tex, 'static': sta}
    for _ in range(2):
        net(x, c, v, e4e_results=e4e)
    rt.profile_begin()
    net(x, c, v, e4e_results=e4e)
    rep = rt.profile_report()
tot = sum(r['ms'] for r in rep.values())
print(f'total {tot:.3f} ms, {sum(r["launches"] for r in rep.values())} launches')
for k, r in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])[:45]:
    print(f'{k:52s} {r["ms"]:8.3f} ms  x {r["launches"]:4d}')
