"""Mint golden vectors of the inversion encoder by executing the UNMODIFIED reference
``encoder_inversion.models.uvnet.inversionNet`` (read-only tree at /root/reference) on CPU.

    python tests/golden/make_golden_encoder.py        # writes tests/golden/encoder.npz     (T=2, 64^2 x 16+16: small, fast)
    python tests/golden/make_golden_encoder.py c3     # writes tests/golden/encoder_c3.npz  (BASELINE configs[2]: T=4, 128^2 x 48+48)
(build container only)

Mode flags follow eval_seq.py:91-97: the whole inversionNet in train mode, then ``input_layer`` / ``body`` of the two
UNets back to eval -- e4e and the UNet decoders normalise with batch statistics.  The two random draws of the T-frame
render inside AR_eval_forward (evaluation=False) are pinned as in make_golden.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference_generator, pinned_draws, state_hash  # noqa: E402  (sets up sys.path + turtle shim)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from fingerprint import fingerprint, pack  # noqa: E402
from invertavatar_b200 import synth  # noqa: E402

T, RES, DC, DF = 2, 64, 16, 16
NAME = 'encoder.npz'
if 'c3' in sys.argv[1:]:      # BASELINE configs[2] at its stated size (train-mode BatchNorm over 4 frames, 48 random-u importance samples)
    T, RES, DC, DF = 4, 128, 48, 48
    NAME = 'encoder_c3.npz'


def build_reference_inversion_net():
    from encoder_inversion.models.uvnet import inversionNet
    generator = build_reference_generator(DC, DF)
    torch.manual_seed(1)
    G = inversionNet(generator=generator, encoding_triplane=True, encoding_texture=True).train().requires_grad_(False)
    synth.randomize_encoder(G)
    G.unet_encoder.triplane_unet.input_layer.eval(); G.unet_encoder.triplane_unet.body.eval()
    G.unet_encoder.texture_unet.input_layer.eval(); G.unet_encoder.texture_unet.body.eval()
    G.generator.neural_rendering_resolution = RES
    return G


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    with torch.no_grad():
        G = build_reference_inversion_net()
        out['enc/state_hash'] = np.frombuffer(state_hash(G.state_dict()).encode(), dtype=np.uint8)
        out['enc/meta'] = np.asarray([T, RES, DC, DF], dtype=np.int64)
        x, c, v = synth.encoder_inputs(T)
        # e4e: train-mode BatchNorm on a single image (eval_seq.py:164) and, for coverage, eval mode
        # (eval first: a train-mode pass updates the running statistics as a side effect)
        G.encoder.eval()
        out['enc/ws_eval'] = G.encode(x['image'][:1]).numpy()
        G.encoder.train()
        ws = G.encode(x['image'][:1])
        out['enc/ws_train'] = ws.numpy()
        tex = G.generator.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        sta = G.generator.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        e4e = {'w': ws, 'texture': tex, 'static': sta}
        r_list = [None, None]
        for call in range(2):   # second call carries the ConvGRU states of the first (eval_seq.py:173-190)
            jit = synth.depth_jitter(T, RES * RES, DC, seed=20 + call)
            u = synth.importance_u(T, RES * RES, DF, seed=30 + call)
            captured = {}
            h1 = G.unet_encoder.texture_unet.register_forward_hook(lambda m, i, o: captured.__setitem__('x_input', i[0]))
            with pinned_draws(jit.clone(), u):
                upd, r_list = G.AR_eval_forward(x, c, v, ws, r_list, e4e_results=e4e, return_fake=False)
            h1.remove()
            tag = f'enc/ar{call}'
            pack(f'{tag}/x_input', fingerprint(captured['x_input']), out)
            for i, t in enumerate(upd['texture']):
                pack(f'{tag}/texture{i}', fingerprint(t), out)
            for i, t in enumerate(upd['static']):
                pack(f'{tag}/static{i}', fingerprint(t), out)
            for n in range(2):
                for i, t in enumerate(r_list[n]):
                    pack(f'{tag}/r{n}_{i}', fingerprint(t), out)
    np.savez_compressed(os.path.join(HERE, NAME), **out)
    print(NAME, len(out))


if __name__ == '__main__':
    main()
