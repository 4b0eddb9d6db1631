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
