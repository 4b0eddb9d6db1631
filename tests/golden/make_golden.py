"""Mint golden vectors by executing the UNMODIFIED reference (read-only tree at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures; these files are what pins the oracle
(``oracle/``) to the reference.  Inputs come from ``invertavatar_b200.synth`` (seeded).  The reference
generator is random-initialised under ``torch.manual_seed(0)`` with the train_avatar_texture.py default architecture
and then ``synth.randomize_noise_and_wavg`` so the noise / truncation paths are exercised; its state-dict hash is
stored so the tests can prove they rebuilt the same weights.

The two random draws of the reference renderer (torch.rand_like, renderer.py:406; torch.rand, renderer.py:453)
are replaced by the supplied tensors through a temporary monkeypatch of the two torch functions; nothing in
the reference tree is modified.
"""
import contextlib
import hashlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.normpath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'

sys.dont_write_bytecode = True
sys.modules.setdefault('turtle', types.SimpleNamespace(update=None))  # triplane_v20.py:12 stray import needs tkinter
sys.path.insert(0, REF)
sys.path.insert(1, REPO)
sys.path.insert(2, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from fingerprint import fingerprint, pack  # noqa: E402
from invertavatar_b200 import synth  # noqa: E402


def state_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@contextlib.contextmanager
def pinned_draws(jitter=None, u=None):
    """Make torch.rand_like / torch.rand return the supplied tensors (once each) while the reference runs."""
    orig_rand_like, orig_rand = torch.rand_like, torch.rand
    state = {'jitter': jitter, 'u': u}

    def rand_like(t, *a, **k):
        j = state['jitter']
        assert j is not None and tuple(j.shape) == tuple(t.shape), (None if j is None else j.shape, t.shape)
        state['jitter'] = None
        return j.to(t.dtype)

    def rand(*size, **k):
        uu = state['u']
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        assert uu is not None and tuple(uu.shape) == shape, (None if uu is None else uu.shape, shape)
        state['u'] = None
        return uu.clone()

    torch.rand_like, torch.rand = rand_like, rand
    try:
        yield
    finally:
        torch.rand_like, torch.rand = orig_rand_like, orig_rand


def build_reference_generator(Dc, Df):
    from training_avatar_texture.triplane_v20 import TriPlaneGenerator
    torch.manual_seed(0)
    G = TriPlaneGenerator(**synth.generator_kwargs(Dc, Df)).eval().requires_grad_(False)
    synth.randomize_noise_and_wavg(G)
    return G


# ------------------------------------------------------------------------------------------------------------
def golden_ops(out):
    """Op-level vectors from the reference's own `_ref` implementations (the CPU branches of torch_utils.ops)."""
    from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, filtered_lrelu
    from training_avatar_texture.networks_stylegan2_new import modulated_conv2d
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 7, 6, generator=g) * 2
    b = torch.randn(5, generator=g)
    out['bias_act/x'], out['bias_act/b'] = x.numpy(), b.numpy()
    for act in bias_act.activation_funcs.keys():
        out[f'bias_act/{act}'] = bias_act.bias_act(x, b, act=act).numpy()
    out['bias_act/lrelu_gain_clamp'] = bias_act.bias_act(x, b, act='lrelu', alpha=0.1, gain=1.7, clamp=1.5).numpy()
    out['bias_act/linear_dim3'] = bias_act.bias_act(x, torch.arange(6.0), dim=3, act='linear').numpy()

    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    x = torch.randn(2, 3, 8, 9, generator=g)
    out['upfirdn2d/x'], out['upfirdn2d/f'] = x.numpy(), f.numpy()
    out['upfirdn2d/up2'] = upfirdn2d.upsample2d(x, f).numpy()
    out['upfirdn2d/down2'] = upfirdn2d.downsample2d(x, f).numpy()
    out['upfirdn2d/filter'] = upfirdn2d.filter2d(x, f).numpy()
    out['upfirdn2d/pad_fir'] = upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4).numpy()
    out['upfirdn2d/up3_down2_pad'] = upfirdn2d.upfirdn2d(x, f, up=3, down=2, padding=[2, 1, 0, 3], flip_filter=True, gain=2).numpy()
    fa = torch.randn(3, 5, generator=g)
    out['upfirdn2d/fa'] = fa.numpy()
    out['upfirdn2d/asym'] = upfirdn2d.upfirdn2d(x, fa, up=[2, 1], down=[1, 2], padding=[1, 2, 2, 1]).numpy()
    out['upfirdn2d/asym_flip'] = upfirdn2d.upfirdn2d(x, fa, up=[2, 1], down=[1, 2], padding=[1, 2, 2, 1], flip_filter=True).numpy()
    f1 = upfirdn2d.setup_filter([1, 2, 4, 6, 6, 4, 2, 1])  # separable (>= 8 taps)
    out['upfirdn2d/f_sep'] = f1.numpy()
    out['upfirdn2d/sep_up2'] = upfirdn2d.upsample2d(x, f1).numpy()
    out['upfirdn2d/negpad'] = upfirdn2d.upfirdn2d(x, f, up=2, padding=[-1, 2, 3, -2]).numpy()

    x = torch.randn(2, 6, 8, 8, generator=g)
    w = torch.randn(4, 6, 3, 3, generator=g)
    s = torch.randn(2, 6, generator=g) + 1
    noise = torch.randn(8, 8, generator=g) * 0.1
    out['modconv/x'], out['modconv/w'], out['modconv/s'], out['modconv/noise'] = x.numpy(), w.numpy(), s.numpy(), noise.numpy()
    out['modconv/same'] = modulated_conv2d(x, w, s, noise=noise, padding=1).numpy()
    out['modconv/same_unfused'] = modulated_conv2d(x, w, s, noise=noise, padding=1, fused_modconv=False).numpy()
    noise2 = torch.randn(16, 16, generator=g) * 0.1
    out['modconv/noise2'] = noise2.numpy()
    out['modconv/up2'] = modulated_conv2d(x, w, s, noise=noise2, up=2, padding=1, resample_filter=f, flip_weight=False).numpy()
    w1 = torch.randn(3, 6, 1, 1, generator=g)
    out['modconv/w1'] = w1.numpy()
    out['modconv/torgb'] = modulated_conv2d(x, w1, s, demodulate=False).numpy()
    out['conv2d_resample/up2'] = conv2d_resample.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=False).numpy()
    out['conv2d_resample/same'] = conv2d_resample.conv2d_resample(x, w, padding=1).numpy()

    x = torch.randn(1, 3, 10, 10, generator=g)
    bb = torch.randn(3, generator=g)
    f12 = upfirdn2d.setup_filter([1, 3, 3, 1], separable=True) if False else torch.tensor([1., 3., 3., 1.]) / 8
    out['filtered_lrelu/x'], out['filtered_lrelu/b'], out['filtered_lrelu/f'] = x.numpy(), bb.numpy(), f12.numpy()
    out['filtered_lrelu/up2_down2'] = filtered_lrelu.filtered_lrelu(x, fu=f12, fd=f12, b=bb, up=2, down=2, padding=3, clamp=0.9, impl='ref').numpy()
    out['filtered_lrelu/plain'] = filtered_lrelu.filtered_lrelu(x, b=bb, impl='ref').numpy()


def golden_fill_mouth(out):
    """cv2.floodFill through the reference's fill_mouth on hand-made masks (hole, soft pixels, open mouth, leak)."""
    from training_avatar_texture.volumetric_rendering.renderer import fill_mouth
    uv = synth.uvcoords_image(2, res=256)
    a = uv[..., 2].unsqueeze(1).clone()
    # sample 1: soft (non-binary) pixels inside the enclosed mouth and a one-pixel leak to the background
    a[1, 0, 160:170, 120:136] = 0.5
    extra = torch.zeros(1, 1, 256, 256)
    extra[0, 0, 40:200, 60:200] = 1.0
    extra[0, 0, 100:120, 100:140] = 0.0          # enclosed hole
    extra[0, 0, 110, 60:100] = 0.0               # channel from the hole to the outside: not enclosed any more
    extra[0, 0, 150:160, 80:90] = 0.25           # soft enclosed region
    a = torch.cat([a, extra], 0)
    full, mouth = fill_mouth(a.clone(), blur_mouth_edge=False)
    out['fill_mouth/alpha'] = a.numpy()
    out['fill_mouth/full'] = full.numpy()
    out['fill_mouth/mouth'] = mouth.numpy()


def golden_renderer(out):
    """ImportanceRenderer_bsMotion + OSGDecoder + RaySampler_zxc on small planes (res 16, Dc=Df=12), both sampling modes."""
    from training_avatar_texture.triplane_v20 import OSGDecoder
    from training_avatar_texture.volumetric_rendering.renderer import ImportanceRenderer_bsMotion
    from training_avatar_texture.volumetric_rendering.ray_sampler import RaySampler_zxc
    torch.manual_seed(5)
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).eval().requires_grad_(False)
    for p in dec.parameters():
        if p.ndim == 1:
            p.copy_(torch.randn_like(p) * 0.1)
    B, res, Dc, Df = 2, 16, 12, 12
    g = torch.Generator().manual_seed(9)
    planes = torch.randn(B, 3, 32, 32, 32, generator=g)
    cam = synth.cameras(B)
    o, d = RaySampler_zxc()(cam[:, :16].view(-1, 4, 4), cam[:, 16:25].view(-1, 3, 3), res)
    out['renderer/cam'], out['renderer/planes'] = cam.numpy(), planes.numpy()
    out['renderer/rays_o'], out['renderer/rays_d'] = o.numpy(), d.numpy()
    for k, v in dec.state_dict().items():
        out[f'renderer/decoder/{k}'] = v.numpy()
    jit = synth.depth_jitter(B, res * res, Dc)
    u = synth.importance_u(B, res * res, Df)
    out['renderer/jitter'], out['renderer/u'] = jit.numpy(), u.numpy()
    opts = synth.rendering_kwargs(Dc, Df)
    R = ImportanceRenderer_bsMotion()
    for name, ev, white in (('eval', True, False), ('rand', False, False), ('eval_white', True, True)):
        o2 = dict(opts, white_back=white)
        with pinned_draws(jit.clone(), None if ev else u):
            rgb, depth, wsum = R(planes, dec, o, d, o2, evaluation=ev)
        out[f'renderer/{name}/rgb'], out[f'renderer/{name}/depth'], out[f'renderer/{name}/wsum'] = rgb.numpy(), depth.numpy(), wsum.numpy()
    o3 = dict(opts, depth_resolution_importance=0)
    with pinned_draws(jit.clone(), None):
        rgb, depth, wsum = R(planes, dec, o, d, o3, evaluation=True)
    out['renderer/coarse_only/rgb'], out['renderer/coarse_only/depth'] = rgb.numpy(), depth.numpy()


def golden_synthesis(out, tag, res, Dc, Df, B, evaluation=True):
    """Whole-generator run; stage boundaries fingerprinted."""
    G = build_reference_generator(Dc, Df)
    out[f'{tag}/state_hash'] = np.frombuffer(state_hash(G.state_dict()).encode(), dtype=np.uint8)
    z = synth.latents(B)
    cond = synth.frontal_camera(B)
    c = synth.cameras(B)
    uv = synth.uvcoords_image(B)
    jit = synth.depth_jitter(B, res * res, Dc)
    u = None if evaluation else synth.importance_u(B, res * res, Df)
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    captured = {}
    orig_rasterize = G.rasterize

    def rasterize(*a, **k):
        r = orig_rasterize(*a, **k)
        captured['rendering_images'], captured['full_alpha'], captured['mouth'] = r
        return r
    G.rasterize = rasterize
    h1 = G.face_backbone.synthesis.register_forward_hook(lambda m, i, o: captured.__setitem__('stitch', o))
    h2 = G.backbone.synthesis.register_forward_hook(lambda m, i, o: captured.__setitem__('static', [t.clone() for t in o]))
    G.neural_rendering_resolution = res
    with pinned_draws(jit.clone(), u):
        o = G.synthesis(ws, c, {'uvcoords_image': uv}, noise_mode='const', evaluation=evaluation, return_featmap=True)
    h1.remove(); h2.remove()
    out[f'{tag}/ws'] = ws.numpy()
    out[f'{tag}/meta'] = np.asarray([res, Dc, Df, B, int(evaluation)], dtype=np.int64)
    pack(f'{tag}/image', fingerprint(o['image']), out)
    pack(f'{tag}/image_raw', fingerprint(o['image_raw']), out)
    pack(f'{tag}/image_depth', fingerprint(o['image_depth']), out)
    pack(f'{tag}/feature_image', fingerprint(o['feature_image']), out)
    pack(f'{tag}/triplane', fingerprint(o['triplane']), out)
    pack(f'{tag}/stitch', fingerprint(captured['stitch']), out)
    pack(f'{tag}/full_alpha', fingerprint(captured['full_alpha']), out)
    for i, t in enumerate(o['texture']):
        pack(f'{tag}/texture{i}', fingerprint(t), out)
    for i, t in enumerate(captured['static']):
        pack(f'{tag}/static{i}', fingerprint(t), out)
    for i, t in enumerate(captured['rendering_images']):
        pack(f'{tag}/rendering_image{i}', fingerprint(t), out)
    return G, ws, o


def golden_with_texture(out, tag, G, ws_all, res, Dc, Df):
    """eval_seq.py per-frame driver: synthesis_withTexture with precomputed feature lists, batch 1, evaluation=False."""
    ws = ws_all[:1]
    tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, noise_mode='const')
    c = synth.cameras(1, first=3)
    uv = synth.uvcoords_image(1, first=3)
    jit = synth.depth_jitter(1, res * res, Dc, seed=8)
    u = synth.importance_u(1, res * res, Df, seed=12)
    with pinned_draws(jit.clone(), u):
        o = G.synthesis_withTexture(ws, tex, c, {'uvcoords_image': uv}, static_feats=[t.clone() for t in sta],
                                    noise_mode='const', evaluation=False)
    pack(f'{tag}/image', fingerprint(o['image']), out)
    pack(f'{tag}/image_raw', fingerprint(o['image_raw']), out)
    pack(f'{tag}/image_depth', fingerprint(o['image_depth']), out)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    want = set(sys.argv[1:]) or {'ops', 'stages', 'c1', 'c2'}
    with torch.no_grad():
        if 'ops' in want:
            out = {}
            golden_ops(out)
            np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)
            print('ops.npz', len(out))
        if 'stages' in want:
            out = {}
            golden_fill_mouth(out)
            golden_renderer(out)
            np.savez_compressed(os.path.join(HERE, 'stages.npz'), **out)
            print('stages.npz', len(out))
        if 'c1' in want:  # BASELINE config 1: 64^2 neural render x 16 depth, batch 1
            out = {}
            G, ws, _ = golden_synthesis(out, 'c1', 64, 16, 16, 1)
            golden_with_texture(out, 'c1_withtex', G, ws, 64, 16, 16)
            np.savez_compressed(os.path.join(HERE, 'synthesis_c1.npz'), **out)
            print('synthesis_c1.npz', len(out))
        if 'c2' in want:  # headline shape: 128^2 x (48+48), two different frames
            out = {}
            G, ws, _ = golden_synthesis(out, 'c2', 128, 48, 48, 2)
            golden_with_texture(out, 'c2_withtex', G, ws, 128, 48, 48)   # eval_seq.py per-frame driver at the headline size
            np.savez_compressed(os.path.join(HERE, 'synthesis_c2.npz'), **out)
            print('synthesis_c2.npz', len(out))


if __name__ == '__main__':
    main()
