"""Debug helper: AR_eval_forward stage errors (product on cuda:0 vs CPU oracle)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from common import build_inversion_net, golden
from invertavatar_b200 import synth
from oracle import encoder as o_enc, stylegan2 as o_sg
g = golden('encoder.npz')
T, res, Dc, Df = [int(v) for v in g['enc/meta']]
cpu = build_inversion_net(Dc, Df, res)
sd = {k: v.clone() for k, v in cpu.state_dict().items()}
n = copy.deepcopy(cpu).to('cuda')
x, c, v = synth.encoder_inputs(T)
ws = torch.from_numpy(g['enc/ws_train'])
gsd = o_sg.sub(sd, 'generator')
with torch.no_grad():
    tex = o_sg.synthesis_network(o_sg.sub(gsd, 'texture_backbone.synthesis'), ws, return_list=True)
    sta = o_sg.synthesis_network(o_sg.sub(gsd, 'backbone.synthesis'), ws, return_list=True)
    jit = synth.depth_jitter(T, res * res, Dc, seed=20); u = synth.importance_u(T, res * res, Df, seed=30)
    want, _ = o_enc.ar_eval_forward(sd, x, c, v['uvcoords_image'], ws, None, cpu.generator.rendering_kwargs, jit, u,
                                    e4e_results={'w': ws, 'texture': tex, 'static': sta}, neural_rendering_resolution=res, stages=True)
    G = n.generator
    xd = {k: t.cuda() for k, t in x.items()}
    texd = [t.cuda() for t in tex]; stad = [t.cuda() for t in sta]
    G.renderer.depth_jitter = jit.cuda(); G.renderer.importance_u = u.cuda()
    upd, fake, r = n.AR_eval_forward(xd, c.cuda(), {k: t.cuda() for k, t in v.items()}, ws.cuda(), [None, None],
                                     e4e_results={'w': ws.cuda(), 'texture': texd, 'static': stad}, return_fake=True)
    print('e4e image err', (fake['e4e'].cpu() - want['e4e_image']).abs().max().item())
    xi = fake['x_input'].cpu()
    for ch in range(7):
        print('x_input ch', ch, (xi[:, ch] - want['x_input'][:, ch]).abs().max().item())
    for i, (a, b) in enumerate(zip(upd['texture'], want['texture'])):
        print('texture', i, (a.cpu() - b).abs().max().item(), b.abs().max().item())
    for i, (a, b) in enumerate(zip(upd['static'], want['static'])):
        print('static', i, (a.cpu() - b).abs().max().item(), b.abs().max().item())
    for i, (a, b) in enumerate(zip(fake and [], [])):
        pass
    for res_k in want['sft']:
        pass
    # replicate the pytest path
    from golden.fingerprint import fingerprint, unpack
    wsd = ws.cuda()
    tex2 = G.texture_backbone.synthesis(wsd, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    sta2 = G.backbone.synthesis(wsd, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    for i, (a, b) in enumerate(zip(tex2, tex)):
        print('product tex', i, (a.cpu() - b).abs().max().item())
    G.renderer.depth_jitter = jit.cuda(); G.renderer.importance_u = u.cuda()
    upd2, fake2, r2 = n.AR_eval_forward(xd, c.cuda(), {k: t.cuda() for k, t in v.items()}, wsd, [None, None],
                                        e4e_results={'w': wsd, 'texture': tex2, 'static': sta2}, return_fake=True)
    xi2 = fake2['x_input']
    print('x_input product-fed vs oracle', (xi2.cpu() - want['x_input']).abs().max().item())
    fp = unpack('enc/ar0/x_input', g)
    got = fingerprint(xi2.unsqueeze(0))
    d = np.abs(got['sub'] - fp['sub'])
    print('fingerprint err', d.max(), int(d.argmax()), got['sub'][d.argmax()], fp['sub'][d.argmax()], 'step', got['step'], fp['step'])
    got3 = fingerprint(want['x_input'].unsqueeze(0))
    print('oracle fingerprint err', np.abs(got3['sub'] - fp['sub']).max())
