"""Shared test helpers: golden loading, generator construction (identical weights on both sides)."""
import hashlib
import os

import numpy as np
import torch

from invertavatar_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def T(a):
    return torch.from_numpy(np.asarray(a))


def state_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


_GEN_CACHE = {}


def build_generator(Dc=48, Df=48):
    """The product module, random-initialised exactly like the reference under manual_seed(0) (the constructors draw
    the same randn sequence; test_oracle_golden checks the state-dict hash against the reference's), on CPU."""
    from invertavatar_b200.triplane import TriPlaneGenerator
    key = (Dc, Df)
    if key not in _GEN_CACHE:
        torch.manual_seed(0)
        G = TriPlaneGenerator(**synth.generator_kwargs(Dc, Df)).eval().requires_grad_(False)
        synth.randomize_noise_and_wavg(G)
        _GEN_CACHE[key] = G
    return _GEN_CACHE[key]


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float('inf') if mse == 0 else 10.0 * np.log10(4.0 / mse)


_ENC_CACHE = {}


def build_inversion_net(Dc=16, Df=16, res=64):
    """The product inversionNet, random-initialised like the reference (tests/golden/make_golden_encoder.py) with the
    mode flags of eval_seq.py:91-97: everything in train mode except the two UNets' input_layer / body.  On CPU."""
    from invertavatar_b200.encoder import inversionNet
    from invertavatar_b200.triplane import TriPlaneGenerator
    key = (Dc, Df, res)
    if key not in _ENC_CACHE:
        torch.manual_seed(0)
        G = TriPlaneGenerator(**synth.generator_kwargs(Dc, Df)).eval().requires_grad_(False)
        synth.randomize_noise_and_wavg(G)
        torch.manual_seed(1)
        net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).train().requires_grad_(False)
        synth.randomize_encoder(net)
        for u in (net.unet_encoder.triplane_unet, net.unet_encoder.texture_unet):
            u.input_layer.eval()
            u.body.eval()
        net.generator.neural_rendering_resolution = res
        _ENC_CACHE[key] = net
    return _ENC_CACHE[key]


# ---- improved one-shot encoder (tests/golden/make_golden_segformer.py) ----------------------------------
MIT_KW = dict(patch_size=4, embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=True,
              depths=[2, 1, 1, 1], sr_ratios=[8, 4, 2, 1], drop_rate=0.0, drop_path_rate=0.1, in_chans=6)
SEG_RES, SEG_DC, SEG_DF = 64, 16, 16
_SEG_CACHE = {}


def segformer_inputs(kind, seed=41):
    """The inputs make_golden_segformer.py fed the reference."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    if kind == 'tb':
        return torch.randn(2, 96, 24, 24, generator=g)
    if kind == 'mit':
        return torch.randn(2, 6, 64, 64, generator=g)
    if kind == 'texdec':
        return (0.5 * torch.randn(2, 7, 256, 256, generator=g)).clamp(-1, 1)
    if kind == 'tridec':
        return (0.5 * torch.randn(2, 6, 256, 256, generator=g)).clamp(-1, 1)
    raise KeyError(kind)


def segformer_forward_draws(B=1):
    return [(synth.depth_jitter(B, SEG_RES * SEG_RES, SEG_DC, seed=50 + i), synth.importance_u(B, SEG_RES * SEG_RES, SEG_DF, seed=60 + i))
            for i in range(2)]


def build_segformer_part(kind):
    """Product modules with the weights of make_golden_segformer.py (name-keyed, so constructor draw order is irrelevant); CPU, eval."""
    from functools import partial
    from invertavatar_b200 import segformer as sf
    if kind not in _SEG_CACHE:
        if kind == 'tb':
            m = sf.transformer_block(in_chans=96, num_vit=2)
        elif kind == 'mit':
            m = sf.MixVisionTransformer(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **MIT_KW)
        elif kind == 'texdec':
            m = sf.TriPlanefeat_SegformerDecoder(inp_ch=7, res=256)
        elif kind == 'tridec':
            m = sf.TriPlaneSFTfeat_SegformerDecoder(inp_ch=6, res=256)
        else:
            raise KeyError(kind)
        _SEG_CACHE[kind] = synth.randomize_by_name(m.eval().requires_grad_(False))
    return _SEG_CACHE[kind]


def build_os_inversion_net():
    """The product uvnet_new.inversionNet as make_golden_segformer.py built the reference's: generator under seed 0, e4e under seed 1
    (+ randomize_encoder), the two decoders by name; eval mode (eval_updated_os.py:93).  On CPU."""
    from invertavatar_b200.segformer import inversionNet
    from invertavatar_b200.triplane import TriPlaneGenerator
    if 'net' not in _SEG_CACHE:
        torch.manual_seed(0)
        G = TriPlaneGenerator(**synth.generator_kwargs(SEG_DC, SEG_DF)).eval().requires_grad_(False)
        synth.randomize_noise_and_wavg(G)
        G.neural_rendering_resolution = SEG_RES
        torch.manual_seed(1)
        net = inversionNet(generator=G, encoding_triplane=True, encoding_texture=True).eval().requires_grad_(False)
        synth.randomize_encoder(net)
        synth.randomize_by_name(net.unet_encoder)
        _SEG_CACHE['net'] = net
    return _SEG_CACHE['net']
