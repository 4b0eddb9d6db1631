"""Mint golden vectors of the "improved one-shot" inversion encoder by executing the UNMODIFIED reference
(``encoder_inversion.models.{uvnet_new, unet_transformer, mmseg.mix_transformer}``, read-only tree at /root/reference) on CPU.

    python tests/golden/make_golden_segformer.py        # writes tests/golden/segformer.npz      (build container only)

``timm`` is a dependency of mmseg/mix_transformer.py:11-13 that this image does not have.  The three helpers the file uses
are supplied by a stub module installed in ``sys.modules`` before the import: ``DropPath`` (identity at inference and for
drop_prob == 0 -- it asserts that it is never asked to drop), ``to_2tuple`` and ``trunc_normal_`` (torch.nn.init's); the two
other imported names (``register_model``, ``_cfg``) are unused by the file.  Nothing in the reference tree is modified.

Weights: constructors run under a seed, then ``synth.randomize_by_name`` overwrites every tensor of the transformer / decoder
modules from a generator keyed by the tensor's state-dict name, so the product modules reproduce them without having to
replay the reference constructors' nested re-initialisation order.
"""
import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference_generator, state_hash  # noqa: E402  (sets up sys.path + turtle shim)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

from fingerprint import fingerprint, pack  # noqa: E402
from invertavatar_b200 import synth  # noqa: E402

RES, DC, DF = 64, 16, 16
MIT_KW = dict(patch_size=4, embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=True,
              depths=[2, 1, 1, 1], sr_ratios=[8, 4, 2, 1], drop_rate=0.0, drop_path_rate=0.1, in_chans=6)


def install_timm_stub():
    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training or self.drop_prob == 0.0, 'stochastic depth only acts in training'
            return x
    mods = {n: types.ModuleType(n) for n in ('timm', 'timm.models', 'timm.models.layers', 'timm.models.registry',
                                             'timm.models.vision_transformer')}
    mods['timm.models.layers'].DropPath = DropPath
    mods['timm.models.layers'].to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    mods['timm.models.layers'].trunc_normal_ = nn.init.trunc_normal_
    mods['timm.models.registry'].register_model = lambda f: f
    mods['timm.models.vision_transformer']._cfg = lambda **k: k
    sys.modules.update(mods)


@contextlib.contextmanager
def pinned_draw_sequence(draws):
    """torch.rand_like / torch.rand return the supplied (jitter, u) pairs, one pair per renderer call, in order."""
    orig_rand_like, orig_rand = torch.rand_like, torch.rand
    jit = [d[0] for d in draws]
    us = [d[1] for d in draws]

    def rand_like(t, *a, **k):
        j = jit.pop(0)
        assert tuple(j.shape) == tuple(t.shape), (j.shape, t.shape)
        return j.to(t.dtype)

    def rand(*size, **k):
        u = us.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        assert tuple(u.shape) == shape, (u.shape, shape)
        return u.clone()
    torch.rand_like, torch.rand = rand_like, rand
    try:
        yield
    finally:
        torch.rand_like, torch.rand = orig_rand_like, orig_rand


def segformer_inputs(kind, seed=41):
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


def forward_draws(B=1):
    return [(synth.depth_jitter(B, RES * RES, DC, seed=50 + i), synth.importance_u(B, RES * RES, DF, seed=60 + i)) for i in range(2)]


def main():
    install_timm_stub()
    from functools import partial
    from encoder_inversion.models import unet_transformer
    from encoder_inversion.models.mmseg import mix_transformer
    from encoder_inversion.models.uvnet_new import inversionNet
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    with torch.no_grad():
        # (1) transformer_block alone (mix_transformer.py:453-472)
        torch.manual_seed(2)
        tb = synth.randomize_by_name(mix_transformer.transformer_block(in_chans=96, num_vit=2).eval())
        out['tb/state_hash'] = np.frombuffer(state_hash(tb.state_dict()).encode(), dtype=np.uint8)
        pack('tb/out', fingerprint(tb(segformer_inputs('tb'))), out)
        # (2) MixVisionTransformer with spatial-reduction attention, q/kv bias, eps 1e-6 (mix_transformer.py:201-376)
        mit = mix_transformer.MixVisionTransformer(norm_layer=partial(nn.LayerNorm, eps=1e-6), **MIT_KW).eval()
        synth.randomize_by_name(mit)
        out['mit/state_hash'] = np.frombuffer(state_hash(mit.state_dict()).encode(), dtype=np.uint8)
        for i, o in enumerate(mit(segformer_inputs('mit'))):
            pack(f'mit/out{i}', fingerprint(o), out)
        # (3) the two decoders, eval mode (eval_updated_os.py:93) and with train-mode BatchNorm in the decoder (batch statistics)
        for kind, cls, ch in (('texdec', unet_transformer.TriPlanefeat_SegformerDecoder, 7), ('tridec', unet_transformer.TriPlaneSFTfeat_SegformerDecoder, 6)):
            m = synth.randomize_by_name(cls(inp_ch=ch, res=256).eval())
            out[f'{kind}/state_hash'] = np.frombuffer(state_hash(m.state_dict()).encode(), dtype=np.uint8)
            x = segformer_inputs(kind)
            for mode in ('eval', 'train'):
                if mode == 'train':
                    m.train()
                    m.input_layer.eval()
                    m.body.eval()
                o = m(x)
                items = enumerate(o) if isinstance(o, list) else o.items()
                for k, t in items:
                    pack(f'{kind}/{mode}/{k}', fingerprint(t), out)
        # (4) uvnet_new.inversionNet.forward end to end, eval mode, one source image
        generator = build_reference_generator(DC, DF)
        generator.neural_rendering_resolution = RES
        torch.manual_seed(1)
        G = inversionNet(generator=generator, encoding_triplane=True, encoding_texture=True).eval().requires_grad_(False)
        synth.randomize_encoder(G)
        synth.randomize_by_name(G.unet_encoder)
        out['fwd/state_hash'] = np.frombuffer(state_hash(G.state_dict()).encode(), dtype=np.uint8)
        out['fwd/meta'] = np.asarray([RES, DC, DF], dtype=np.int64)
        x, c, v = synth.encoder_inputs(1)
        with pinned_draw_sequence(forward_draws(1)):
            o = G(x, c, v, return_feats=True, visualize_input=True)
        out['fwd/w'] = o['w'].numpy()
        for k in ('image', 'e4e_image', 'x_input', 'image_raw', 'image_depth'):
            pack(f'fwd/{k}', fingerprint(o[k]), out)
        for i, t in enumerate(o['texture']):
            pack(f'fwd/texture{i}', fingerprint(t), out)
        for i, t in enumerate(o['static']):
            pack(f'fwd/static{i}', fingerprint(t), out)
    np.savez_compressed(os.path.join(HERE, 'segformer.npz'), **out)
    print('segformer.npz', len(out))


if __name__ == '__main__':
    main()
