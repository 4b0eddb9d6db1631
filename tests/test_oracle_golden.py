"""CPU: the oracle (oracle/) against golden vectors minted from the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; tolerances are the fp32 reassociation floor."""
import numpy as np
import pytest
import torch

from common import T, build_generator, golden, state_hash
from fingerprint import compare, unpack
from invertavatar_b200 import synth
from oracle import ops as o_ops
from oracle import renderer as o_rd
from oracle import stylegan2 as o_sg
from oracle import triplane as o_tp

ATOL = 2e-5


def close(a, b, atol=ATOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) if a.size else 0.0
    assert err <= atol, f'max-abs {err:.3e} > {atol:.1e}'


def test_bias_act_golden():
    g = golden('ops.npz')
    x, b = T(g['bias_act/x']), T(g['bias_act/b'])
    for act in o_ops.ACT_DEFAULTS:
        close(o_ops.bias_act(x, b, act=act).numpy(), g[f'bias_act/{act}'])
    close(o_ops.bias_act(x, b, act='lrelu', alpha=0.1, gain=1.7, clamp=1.5).numpy(), g['bias_act/lrelu_gain_clamp'])
    close(o_ops.bias_act(x, torch.arange(6.0), dim=3).numpy(), g['bias_act/linear_dim3'])


def test_upfirdn2d_golden():
    g = golden('ops.npz')
    x, f = T(g['upfirdn2d/x']), T(g['upfirdn2d/f'])
    close(o_ops.upsample2d(x, f).numpy(), g['upfirdn2d/up2'])
    close(o_ops.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4).numpy(), g['upfirdn2d/pad_fir'])
    close(o_ops.upfirdn2d(x, f, up=3, down=2, padding=[2, 1, 0, 3], flip_filter=True, gain=2).numpy(), g['upfirdn2d/up3_down2_pad'])
    close(o_ops.upfirdn2d(x, f, up=2, padding=[-1, 2, 3, -2]).numpy(), g['upfirdn2d/negpad'])
    close(o_ops.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]).numpy(), g['upfirdn2d/down2'])
    close(o_ops.upfirdn2d(x, f, padding=[2, 1, 2, 1]).numpy(), g['upfirdn2d/filter'])


def test_modconv_golden():
    g = golden('ops.npz')
    x, w, s = T(g['modconv/x']), T(g['modconv/w']), T(g['modconv/s'])
    f = T(g['upfirdn2d/f'])
    close(o_sg.modulated_conv2d(x, w, s, noise=T(g['modconv/noise']), padding=1).numpy(), g['modconv/same'])
    close(o_sg.modulated_conv2d(x, w, s, noise=T(g['modconv/noise']), padding=1, fused_modconv=False).numpy(), g['modconv/same_unfused'])
    close(o_sg.modulated_conv2d(x, w, s, noise=T(g['modconv/noise2']), up=2, padding=1, resample_filter=f, flip_weight=False).numpy(),
          g['modconv/up2'])
    close(o_sg.modulated_conv2d(x, T(g['modconv/w1']), s, demodulate=False).numpy(), g['modconv/torgb'])
    close(o_ops.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=False).numpy(), g['conv2d_resample/up2'])
    close(o_ops.conv2d_resample(x, w, padding=1).numpy(), g['conv2d_resample/same'])


def test_filtered_lrelu_golden():
    g = golden('ops.npz')
    x, b, f = T(g['filtered_lrelu/x']), T(g['filtered_lrelu/b']), T(g['filtered_lrelu/f'])
    close(o_ops.filtered_lrelu(x, fu=f, fd=f, b=b, up=2, down=2, padding=3, clamp=0.9).numpy(), g['filtered_lrelu/up2_down2'])
    close(o_ops.filtered_lrelu(x, b=b).numpy(), g['filtered_lrelu/plain'])


def test_fill_mouth_golden():
    g = golden('stages.npz')
    full, mouth = o_tp.fill_mouth(T(g['fill_mouth/alpha']).clone())
    close(full.numpy(), g['fill_mouth/full'], 0)
    close(mouth.numpy(), g['fill_mouth/mouth'], 0)
    assert float(T(g['fill_mouth/mouth'])[0].sum()) > 50      # the synthetic mouth hole is really filled
    assert float(T(g['fill_mouth/mouth'])[2, 0, 100:120, 100:140].sum()) == 0   # the leaking hole is not


def _decoder_sd(g):
    return {k[len('renderer/decoder/'):]: T(g[k]) for k in g.files if k.startswith('renderer/decoder/')}


def test_ray_sampler_golden():
    g = golden('stages.npz')
    cam = T(g['renderer/cam'])
    o, d = o_rd.ray_sampler_zxc(cam[:, :16].view(-1, 4, 4), cam[:, 16:25].view(-1, 3, 3), 16)
    close(o.numpy(), g['renderer/rays_o'], 1e-6)
    close(d.numpy(), g['renderer/rays_d'], 1e-6)


@pytest.mark.parametrize('name,ev,white', [('eval', True, False), ('rand', False, False), ('eval_white', True, True)])
def test_renderer_golden(name, ev, white):
    g = golden('stages.npz')
    opts = dict(synth.rendering_kwargs(12, 12), white_back=white)
    rgb, depth, wsum = o_rd.importance_renderer(_decoder_sd(g), T(g['renderer/planes']), T(g['renderer/rays_o']), T(g['renderer/rays_d']),
                                                opts, T(g['renderer/jitter']), evaluation=ev, u=None if ev else T(g['renderer/u']))
    close(rgb.numpy(), g[f'renderer/{name}/rgb'])
    close(depth.numpy(), g[f'renderer/{name}/depth'])
    close(wsum.numpy(), g[f'renderer/{name}/wsum'])


def test_renderer_coarse_only_golden():
    g = golden('stages.npz')
    opts = synth.rendering_kwargs(12, 0)
    rgb, depth, _ = o_rd.importance_renderer(_decoder_sd(g), T(g['renderer/planes']), T(g['renderer/rays_o']), T(g['renderer/rays_d']),
                                             opts, T(g['renderer/jitter']), evaluation=True)
    close(rgb.numpy(), g['renderer/coarse_only/rgb'])
    close(depth.numpy(), g['renderer/coarse_only/depth'])


def _oracle_synthesis(tag, npz):
    g = golden(npz)
    res, Dc, Df, B, ev = [int(v) for v in g[f'{tag}/meta']]
    G = build_generator(Dc, Df)
    sd = G.state_dict()
    assert state_hash(sd) == bytes(g[f'{tag}/state_hash']).decode(), 'module init does not reproduce the reference weights'
    z, cond, c, uv = synth.latents(B), synth.frontal_camera(B), synth.cameras(B), synth.uvcoords_image(B)
    ws = o_tp.mapping(sd, z, cond, G.rendering_kwargs, truncation_psi=0.7, truncation_cutoff=14)
    close(ws.numpy(), g[f'{tag}/ws'], 1e-5)
    jit = synth.depth_jitter(B, res * res, Dc)
    out = o_tp.synthesis(sd, ws, c, uv, G.rendering_kwargs, jit, evaluation=bool(ev), neural_rendering_resolution=res, stages=True)
    return g, G, sd, ws, out


def _check_stages(g, tag, out, atol):
    for i, t in enumerate(out['texture_feats']):
        compare(t, unpack(f'{tag}/texture{i}', g), atol, f'texture{i}')
    for i, t in enumerate(out['static_feats']):
        compare(t, unpack(f'{tag}/static{i}', g), atol, f'static{i}')
    for i, t in enumerate(out['rendering_images']):
        compare(t, unpack(f'{tag}/rendering_image{i}', g), atol, f'rendering_image{i}')
    compare(out['full_alpha'], unpack(f'{tag}/full_alpha', g), atol, 'full_alpha')
    compare(out['rendering_stitch'], unpack(f'{tag}/stitch', g), atol, 'stitch')
    compare(out['triplane'], unpack(f'{tag}/triplane', g), atol, 'triplane')
    compare(out['feature_image'], unpack(f'{tag}/feature_image', g), atol, 'feature_image')
    compare(out['image_raw'], unpack(f'{tag}/image_raw', g), atol, 'image_raw')
    compare(out['image_depth'], unpack(f'{tag}/image_depth', g), atol, 'image_depth')
    compare(out['image'], unpack(f'{tag}/image', g), atol, 'image')


def test_synthesis_c1_golden():
    """BASELINE config 1: 64^2 x (16+16), batch 1, CPU only."""
    g, G, sd, ws, out = _oracle_synthesis('c1', 'synthesis_c1.npz')
    _check_stages(g, 'c1', out, 5e-5)
    # eval_seq.py per-frame driver (synthesis_withTexture, evaluation=False -> random importance u)
    ws1 = ws[:1]
    tex = o_sg.synthesis_network(o_sg.sub(sd, 'texture_backbone.synthesis'), ws1, return_list=True)
    sta = o_sg.synthesis_network(o_sg.sub(sd, 'backbone.synthesis'), ws1, return_list=True)
    o = o_tp.synthesis_with_texture(sd, ws1, tex, synth.cameras(1, first=3), synth.uvcoords_image(1, first=3), G.rendering_kwargs,
                                    synth.depth_jitter(1, 64 * 64, 16, seed=8), static_feats=sta, evaluation=False,
                                    u=synth.importance_u(1, 64 * 64, 16, seed=12), neural_rendering_resolution=64)
    for k in ('image', 'image_raw', 'image_depth'):
        compare(o[k], unpack(f'c1_withtex/{k}', g), 5e-5, 'withtex/' + k)


def test_synthesis_c2_golden():
    """Headline shape: 128^2 x (48+48), two frames with different latents / cameras / UV conditions."""
    g, G, sd, ws, out = _oracle_synthesis('c2', 'synthesis_c2.npz')
    _check_stages(g, 'c2', out, 5e-5)
    # eval_seq.py per-frame driver at the headline size (synthesis_withTexture, evaluation=False -> 48 random-u importance samples)
    ws1 = ws[:1]
    tex = o_sg.synthesis_network(o_sg.sub(sd, 'texture_backbone.synthesis'), ws1, return_list=True)
    sta = o_sg.synthesis_network(o_sg.sub(sd, 'backbone.synthesis'), ws1, return_list=True)
    o = o_tp.synthesis_with_texture(sd, ws1, tex, synth.cameras(1, first=3), synth.uvcoords_image(1, first=3), G.rendering_kwargs,
                                    synth.depth_jitter(1, 128 * 128, 48, seed=8), static_feats=sta, evaluation=False,
                                    u=synth.importance_u(1, 128 * 128, 48, seed=12), neural_rendering_resolution=128)
    for k in ('image', 'image_raw', 'image_depth'):
        compare(o[k], unpack(f'c2_withtex/{k}', g), 5e-5, 'c2_withtex/' + k)
