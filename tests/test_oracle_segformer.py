"""CPU: the oracle of the improved one-shot encoder (oracle/segformer.py) against golden vectors minted from the
unmodified reference (tests/golden/make_golden_segformer.py), and the product modules' state-dicts against the reference's."""
import numpy as np
import pytest
import torch

from common import (MIT_KW, SEG_RES, build_os_inversion_net, build_segformer_part, golden, segformer_forward_draws, segformer_inputs,
                    state_hash)
from golden.fingerprint import compare, unpack
from invertavatar_b200 import synth
from oracle import segformer as o_sf

ATOL = 2e-4     # relative to the tensor's magnitude where that exceeds 1 (fp32 reassociation; random weights of unit gain)


def _cmp(t, g, key, atol=ATOL):
    fp = unpack(key, g)
    scale = max(1.0, float(np.abs(fp['sub']).max()))
    return compare(t, fp, atol * scale, key)


@pytest.mark.parametrize('kind', ['tb', 'mit', 'texdec', 'tridec'])
def test_state_dict_matches_reference(kind):
    g = golden('segformer.npz')
    assert state_hash(build_segformer_part(kind).state_dict()) == bytes(g[f'{kind}/state_hash']).decode(), \
        f'{kind}: state-dict names / order / shapes differ from the reference module'


def test_inversion_net_state_dict_matches_reference():
    g = golden('segformer.npz')
    assert state_hash(build_os_inversion_net().state_dict()) == bytes(g['fwd/state_hash']).decode()


def test_transformer_block_golden():
    g = golden('segformer.npz')
    with torch.no_grad():
        out = o_sf.transformer_block(build_segformer_part('tb').state_dict(), segformer_inputs('tb'), 2)
    _cmp(out, g, 'tb/out')


def test_mix_vision_transformer_golden():
    g = golden('segformer.npz')
    with torch.no_grad():
        outs = o_sf.mix_vision_transformer(build_segformer_part('mit').state_dict(), segformer_inputs('mit'), MIT_KW['depths'],
                                           MIT_KW['num_heads'], MIT_KW['sr_ratios'], eps=1e-6)
    for i, o in enumerate(outs):
        _cmp(o, g, f'mit/out{i}')


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_decoders_golden(mode):
    g = golden('segformer.npz')
    with torch.no_grad():
        outs = o_sf.texture_segformer_decoder(build_segformer_part('texdec').state_dict(), segformer_inputs('texdec'), False, mode == 'train')
        for i, o in enumerate(outs):
            _cmp(o, g, f'texdec/{mode}/{i}')
        outd = o_sf.triplane_segformer_decoder(build_segformer_part('tridec').state_dict(), segformer_inputs('tridec'), False, mode == 'train')
        for res, o in outd.items():
            _cmp(o, g, f'tridec/{mode}/{res}')


def test_forward_golden():
    """uvnet_new.inversionNet.forward end to end (eval mode, two pinned renders)."""
    g = golden('segformer.npz')
    net = build_os_inversion_net()
    x, c, v = synth.encoder_inputs(1)
    with torch.no_grad():
        out = o_sf.forward(net.state_dict(), x, c, v['uvcoords_image'], net.generator.rendering_kwargs, segformer_forward_draws(1),
                           training=False, neural_rendering_resolution=SEG_RES)
    assert np.abs(out['w'].numpy() - g['fwd/w']).max() <= ATOL * max(1.0, float(np.abs(g['fwd/w']).max()))
    for k in ('x_input', 'e4e_image', 'image', 'image_raw', 'image_depth'):
        _cmp(out[k].clamp(-1, 1) if k == 'x_input' else out[k], g, f'fwd/{k}')
    for i, t in enumerate(out['texture']):
        _cmp(t, g, f'fwd/texture{i}')
    for i, t in enumerate(out['static']):
        _cmp(t, g, f'fwd/static{i}')
