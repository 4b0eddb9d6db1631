"""eval_updated_os.py one-shot path: e4e encode + uvnet_new.inversionNet.forward (two 128^2 x 48+48 renders, the two
SegFormer-style decoders) on one source image, then the per-frame synthesis_withTexture driver, on cuda:0, with the per-kernel
device-time breakdown and the roofline position of the two dominant kernels (tcgen05 GEMMs / mma.sync flash attention).
Prints one JSON line.   python tools/bench_oneshot.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from invertavatar_b200 import runtime as rt  # noqa: E402
from invertavatar_b200 import synth  # noqa: E402
from invertavatar_b200.segformer import inversionNet  # noqa: E402
from invertavatar_b200.triplane import TriPlaneGenerator  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
x = {k: t.to(dev) for k, t in x.items()}
c = c.to(dev)
v = {k: t.to(dev) for k, t in v.items()}


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def attention_flops(net):
    """2*MAC of q k^T and p v over every Block of the two decoders at 256^2 input, batch 1 (tokens: 8^2, 16^2, 32^2, 64^2)."""
    total = 0
    for dec in (net.unet_encoder.texture_unet, net.unet_encoder.triplane_unet):
        for up, n_tok in ((dec.up1, 64), (dec.up2, 256), (dec.up3, 1024), (dec.up4, 4096)):
            total += len(up.transformer.ViT) * 2 * 2 * n_tok * n_tok * 1024
    return total


with torch.no_grad():
    ms_enc, ws = timed(lambda: net.encode(x['image']), steps)
    tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
    e4e = {'w': ws, 'texture': tex, 'static': sta}
    ms_fwd, out = timed(lambda: net(x, c, v, e4e_results=e4e, return_feats=True), steps)
    # the same forward replayed as one CUDA graph (the eager forward is ~1240 launches of ~25 us of host time each)
    from invertavatar_b200.graphs import GraphedCall
    gin = dict(image=x['image'], uv=x['uv'], c=c, uvc=v['uvcoords_image'], w=ws)
    gin.update({f'tex{i}': t for i, t in enumerate(tex)})
    gin.update({f'sta{i}': t for i, t in enumerate(sta)})

    def fwd(image, uv, c, uvc, w, **feats):
        e = {'w': w, 'texture': [feats[f'tex{i}'] for i in range(len(tex))], 'static': [feats[f'sta{i}'] for i in range(len(sta))]}
        o = net({'image': image, 'uv': uv}, c, {'uvcoords_image': uvc}, e4e_results=e, return_feats=True)
        return o['image']
    gc = GraphedCall(fwd, gin)
    ms_fwd_graph, _ = timed(lambda: gc(), steps)
    static = e4e['static'][:-1] + out['static'][-1:]        # eval_updated_os.py:172
    ms_frame, img = timed(lambda: net.generator.synthesis_withTexture(ws, out['texture'], c, v, noise_mode='const', static_feats=static,
                                                                      evaluation=True)['image'], 8)
    rt.flop_count_begin()
    rt.profile_begin()
    net(x, c, v, e4e_results=e4e, return_feats=True)
    rep = rt.profile_report()
    fl = rt.flop_count_end()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {}
tot = sum(r['ms'] for r in rep.values())
top = {k: {'ms': round(r['ms'], 3), 'launches': r['launches']} for k, r in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])[:12]}
conv_ms = sum(r['ms'] for k, r in rep.items() if k.startswith('ia_conv_tc'))
att_ms = sum(r['ms'] for k, r in rep.items() if k.startswith('ia_attention'))
att_fl = attention_flops(net)
print(json.dumps({
    'config': 'eval_updated_os.py one-shot path: uvnet_new.inversionNet.forward on one 512^2 source (2 x 128^2 x 48+48 renders + '
              'TriPlanefeat_/TriPlaneSFTfeat_SegformerDecoder at 256^2), then per-frame synthesis_withTexture; 1xB200, random-init weights',
    'encode_ms': ms_enc, 'forward_ms': ms_fwd, 'forward_graph_replay_ms': ms_fwd_graph, 'synthesis_withTexture_ms_per_frame': ms_frame, 'driven_frames_per_s': 1000.0 / ms_frame,
    'device_ms_forward_serialised': tot, 'launches_forward': sum(r['launches'] for r in rep.values()), 'top_kernels': top,
    'gemm': {'kernel': 'conv_tc2_kernel / conv_tc_kernel (every nn.Linear / patch embedding / convolution of the forward)', 'ms': conv_ms,
             'algorithmic_gflop': fl['algorithmic'] / 1e9, 'issued_mma_gflop': fl['issued_mma'] / 1e9,
             'achieved_tflops': fl['algorithmic'] / conv_ms / 1e9 if conv_ms else None,
             'peak_tflops_sustained': peaks.get('bf16_tflops_sustained'), 'bound': 'tensor'},
    'attention': {'kernel': 'attention_tc_kernel (flash attention on mma.sync.m16n8k16, 3-term bf16 split for q k^T and p v; IA_ATTENTION=simt: '
                            'attention_kernel<256>, fp32 CUDA cores)', 'ms': att_ms, 'algorithmic_gflop': att_fl / 1e9,
                  'achieved_tflops': att_fl / att_ms / 1e9 if att_ms else None,
                  'issued_mma_tflops': 3 * att_fl / att_ms / 1e9 if att_ms else None, 'bound': 'tensor (mma.sync) / shared-memory operand fetch'},
}))
