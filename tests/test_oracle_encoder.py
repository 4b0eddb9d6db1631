"""CPU: the encoder oracle (oracle/encoder.py) against golden vectors minted from the unmodified reference
(tests/golden/make_golden_encoder.py), and the product modules' state-dict against the reference's."""
import numpy as np
import pytest
import torch

from common import build_inversion_net, golden, state_hash
from golden.fingerprint import compare, unpack
from invertavatar_b200 import synth
from oracle import encoder as o_enc
from oracle import stylegan2 as o_sg

ATOL = 2e-4   # fp32 reassociation across ~60 stacked convolutions (the reference itself moves by this much between thread counts)


# encoder.npz: T=2, 64^2 x 16+16 (small); encoder_c3.npz: BASELINE configs[2] at its stated size (T=4, 128^2 x 48+48)
NPZ = ['encoder.npz', 'encoder_c3.npz']


def _setup(npz='encoder.npz'):
    g = golden(npz)
    T, res, Dc, Df = [int(v) for v in g['enc/meta']]
    net = build_inversion_net(Dc, Df, res)
    return g, net, T, res, Dc, Df


def test_state_dict_matches_reference():
    g, net, *_ = _setup()
    assert state_hash(net.state_dict()) == bytes(g['enc/state_hash']).decode(), \
        'inversionNet construction does not reproduce the reference parameters (names, order or values differ)'


@pytest.mark.parametrize('npz', NPZ)
def test_encode_golden(npz):
    g, net, T, res, Dc, Df = _setup(npz)
    sd = net.state_dict()
    x, c, v = synth.encoder_inputs(T)
    with torch.no_grad():
        ws_train = o_enc.encode(sd, x['image'][:1], training=True)
        ws_eval = o_enc.encode(sd, x['image'][:1], training=False)
    scale = float(np.abs(g['enc/ws_train']).max())
    assert np.abs(ws_train.numpy() - g['enc/ws_train']).max() <= ATOL * max(1.0, scale)
    assert np.abs(ws_eval.numpy() - g['enc/ws_eval']).max() <= ATOL * max(1.0, float(np.abs(g['enc/ws_eval']).max()))


@pytest.mark.parametrize('npz', NPZ)
def test_ar_eval_forward_golden(npz):
    g, net, T, res, Dc, Df = _setup(npz)
    sd = net.state_dict()
    x, c, v = synth.encoder_inputs(T)
    ws = torch.from_numpy(g['enc/ws_train'])
    gsd = o_sg.sub(sd, 'generator')
    with torch.no_grad():
        tex = o_sg.synthesis_network(o_sg.sub(gsd, 'texture_backbone.synthesis'), ws, return_list=True)
        sta = o_sg.synthesis_network(o_sg.sub(gsd, 'backbone.synthesis'), ws, return_list=True)
        e4e = {'w': ws, 'texture': tex, 'static': sta}
        r_list = [None, None]
        for call in range(2):
            jit = synth.depth_jitter(T, res * res, Dc, seed=20 + call)
            u = synth.importance_u(T, res * res, Df, seed=30 + call)
            upd, r_list = o_enc.ar_eval_forward(sd, x, c, v['uvcoords_image'], ws, r_list, net.generator.rendering_kwargs, jit, u,
                                                e4e_results=e4e, neural_rendering_resolution=res, stages=True)
            tag = f'enc/ar{call}'
            compare(upd['x_input'].unsqueeze(0), unpack(f'{tag}/x_input', g), ATOL, 'x_input')
            for i, t in enumerate(upd['texture']):
                compare(t, unpack(f'{tag}/texture{i}', g), ATOL * 5, f'texture{i}')
            for i, t in enumerate(upd['static']):
                compare(t, unpack(f'{tag}/static{i}', g), ATOL * 5, f'static{i}')
            for n in range(2):
                for i, t in enumerate(r_list[n]):
                    compare(t, unpack(f'{tag}/r{n}_{i}', g), ATOL, f'r{n}_{i}')
