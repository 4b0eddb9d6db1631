"""Debug helper: the pytest path of AR_eval_forward (product-computed e4e features) for compute-sanitizer."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from common import build_inversion_net, golden
from invertavatar_b200 import synth
g = golden('encoder.npz')
T, res, Dc, Df = [int(v) for v in g['enc/meta']]
n = copy.deepcopy(build_inversion_net(Dc, Df, res)).to('cuda')
x, c, v = synth.encoder_inputs(T)
ws = torch.from_numpy(g['enc/ws_train']).cuda()
G = n.generator
with torch.no_grad():
    xd = {k: t.cuda() for k, t in x.items()}
    tex2 = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    sta2 = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    torch.cuda.synchronize(); print('feats ok', flush=True)
    G.renderer.depth_jitter = synth.depth_jitter(T, res * res, Dc, seed=20).cuda(); G.renderer.importance_u = synth.importance_u(T, res * res, Df, seed=30).cuda()
    upd2, r2 = n.AR_eval_forward(xd, c.cuda(), {k: t.cuda() for k, t in v.items()}, ws, [None, None],
                                 e4e_results={'w': ws, 'texture': tex2, 'static': sta2}, return_fake=False)
    torch.cuda.synchronize(); print('AR ok', flush=True)
