"""GPU parity of the fused volume renderer (ia_render) against reference-minted vectors (stages.npz) and the oracle."""
import numpy as np
import pytest
import torch

from common import T, golden
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import synth
from invertavatar_b200.rendering import ImportanceRenderer_bsMotion, RaySampler_zxc
from invertavatar_b200.triplane import OSGDecoder
from oracle import renderer as o_rd

pytestmark = pytest.mark.gpu
DEV = 'cuda'
ATOL = 2e-5     # fp32 path; fast-math exp/log in the decoder and a different summation order


def maxerr(a, b):
    a = a.detach().cpu().double() if hasattr(a, 'detach') else torch.as_tensor(a).double()
    b = b.detach().cpu().double() if hasattr(b, 'detach') else torch.as_tensor(b).double()
    assert tuple(a.shape) == tuple(b.shape), (a.shape, b.shape)
    return float((a - b).abs().max())


def _decoder(g):
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).requires_grad_(False)
    dec.load_state_dict({k[len('renderer/decoder/'):]: T(g[k]) for k in g.files if k.startswith('renderer/decoder/')})
    return dec.to(DEV)


def test_ray_sampler_golden():
    g = golden('stages.npz')
    cam = T(g['renderer/cam']).to(DEV)
    o, d = RaySampler_zxc()(cam[:, :16].view(-1, 4, 4), cam[:, 16:25].view(-1, 3, 3), 16)
    assert maxerr(o, g['renderer/rays_o']) <= 1e-6
    assert maxerr(d, g['renderer/rays_d']) <= 2e-6


@pytest.mark.parametrize('name,ev,white,Df', [('eval', True, False, 12), ('rand', False, False, 12), ('eval_white', True, True, 12),
                                              ('coarse_only', True, False, 0)])
def test_renderer_golden(name, ev, white, Df):
    """Reference API: forward(planes [B,3,32,H,W], decoder, rays_o, rays_d, options, evaluation)."""
    g = golden('stages.npz')
    R = ImportanceRenderer_bsMotion()
    R.depth_jitter = T(g['renderer/jitter']).to(DEV)
    if not ev:
        R.importance_u = T(g['renderer/u']).to(DEV)
    opts = dict(synth.rendering_kwargs(12, Df), white_back=white)
    rgb, depth, wsum = R(T(g['renderer/planes']).to(DEV), _decoder(g), T(g['renderer/rays_o']).to(DEV), T(g['renderer/rays_d']).to(DEV),
                         opts, evaluation=ev)
    assert maxerr(rgb, g[f'renderer/{name}/rgb']) <= ATOL
    assert maxerr(depth, g[f'renderer/{name}/depth']) <= ATOL
    if Df:
        assert maxerr(wsum, g[f'renderer/{name}/wsum']) <= ATOL


def test_renderer_from_camera_matches_explicit_rays():
    """Engine entry (rays generated in-kernel from the camera) == reference API entry (explicit rays)."""
    g = golden('stages.npz')
    planes = T(g['renderer/planes']).to(DEV)
    B = planes.shape[0]
    planes_nhwc = rt.to_nhwc(planes.reshape(B, 96, 32, 32))
    R = ImportanceRenderer_bsMotion()
    R.depth_jitter = T(g['renderer/jitter']).to(DEV)
    feat, depth, wsum = R.render_nhwc(planes_nhwc, _decoder(g), T(g['renderer/cam']).to(DEV), 16, synth.rendering_kwargs(12, 12), evaluation=True)
    assert maxerr(feat.reshape(B, 256, 32), g['renderer/eval/rgb']) <= ATOL
    assert maxerr(depth.reshape(B, 256, 1), g['renderer/eval/depth']) <= ATOL


@pytest.mark.parametrize('res,Dc,Df', [(32, 48, 48), (24, 16, 16), (16, 96, 96), (8, 5, 3)])
def test_renderer_vs_oracle_sweep(res, Dc, Df):
    """Depth-sample / resolution sweep (BASELINE config 5 at reduced ray counts) against the oracle, both sampling modes."""
    gen = torch.Generator().manual_seed(res + Dc)
    B = 2
    planes = torch.randn(B, 3, 32, 48, 48, generator=gen)
    cam = synth.cameras(B, first=5)
    torch.manual_seed(2)
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).requires_grad_(False)
    dec.net[0].bias.copy_(torch.randn(64) * 0.1)
    o, d = o_rd.ray_sampler_zxc(cam[:, :16].view(-1, 4, 4), cam[:, 16:25].view(-1, 3, 3), res)
    jit = synth.depth_jitter(B, res * res, Dc)
    u = synth.importance_u(B, res * res, Df)
    opts = synth.rendering_kwargs(Dc, Df)
    dsd = {k[len('net.'):] if False else k: v for k, v in dec.state_dict().items()}
    for ev in (True, False):
        rgb, depth, wsum = o_rd.importance_renderer(dsd, planes, o, d, opts, jit, evaluation=ev, u=None if ev else u)
        R = ImportanceRenderer_bsMotion()
        R.depth_jitter, R.importance_u = jit.to(DEV), (None if ev else u.to(DEV))
        feat, dd, ww = R.render_nhwc(rt.to_nhwc(planes.reshape(B, 96, 48, 48).to(DEV)), dec.to(DEV), cam.to(DEV), res, opts, evaluation=ev)
        assert maxerr(feat.reshape(B, -1, 32), rgb) <= ATOL, (res, Dc, Df, ev)
        assert maxerr(dd.reshape(B, -1, 1), depth) <= ATOL
        assert maxerr(ww.reshape(B, -1, 1), wsum) <= ATOL
        dec = dec.cpu()


def test_renderer_properties_full_size():
    """Size-independent properties at the headline ray count (128^2 x 48+48, batch 2): weights sum in [0,1], colours in the
    sigmoid range, depth inside [near, far], and a constant radiance field composites to that constant times the weight sum."""
    B, res, Dc, Df = 2, 128, 48, 48
    planes = torch.zeros(B, 256, 256, 96, device=DEV)
    torch.manual_seed(0)
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).requires_grad_(False).to(DEV)
    dec.net[2].bias.copy_(torch.linspace(-1, 1, 33))
    cam = synth.cameras(B).to(DEV)
    R = ImportanceRenderer_bsMotion()
    feat, depth, wsum = R.render_nhwc(planes, dec, cam, res, synth.rendering_kwargs(Dc, Df), evaluation=True)
    assert float(wsum.min()) >= 0 and float(wsum.max()) <= 1 + 1e-5
    # zero planes -> every sample has the same colour c = sigmoid(MLP(0))*1.002-0.001 -> rgb = (c*wsum)*2-1
    x = torch.nn.functional.softplus(dec.net[0].bias)
    y = (dec.net[2].weight / 8.0) @ x + dec.net[2].bias
    c = torch.sigmoid(y[1:]) * 1.002 - 0.001
    expect = (wsum.unsqueeze(-1) * c) * 2 - 1
    assert float((feat - expect).abs().max()) <= 1e-5
    near = 2.7 - 0.45
    assert float(depth.min()) >= near - 1e-4 and float(depth.max()) <= 2.7 + 0.6 + 1e-4


@pytest.mark.parametrize('white', [False, True])
def test_ray_marcher_standalone_vs_oracle(white):
    """MipRayMarcher2 on caller-provided sorted samples (ray_marcher.py:25-57) -- the module-level API; inside the generator the
    marcher is fused into the render kernel."""
    from invertavatar_b200.rendering import MipRayMarcher2
    g = torch.Generator().manual_seed(11)
    B, R, S, C = 2, 333, 24, 32
    depths = (2.2 + torch.rand(B, R, S, 1, generator=g)).sort(dim=2).values
    colors = torch.rand(B, R, S, C, generator=g)
    dens = torch.randn(B, R, S, 1, generator=g) * 3
    dens[0, :7] = -30.0            # empty rays: sum of weights ~ 0 -> depth NaN/inf -> clamped to the global range
    rgb, depth, w = MipRayMarcher2()(colors.to(DEV), dens.to(DEV), depths.to(DEV), {'clamp_mode': 'softplus', 'white_back': white})
    r_rgb, r_depth, r_w = o_rd.ray_march(colors, dens, depths, white_back=white)
    assert tuple(rgb.shape) == (B, R, C) and tuple(depth.shape) == (B, R, 1) and tuple(w.shape) == (B, R, S - 1, 1)
    assert maxerr(w, r_w) <= 2e-6 and maxerr(rgb, r_rgb) <= 1e-5 and maxerr(depth, r_depth) <= 1e-5
