"""CPU: the mesh-condition oracle (oracle/faceverse.py) on the synthetic model -- internal consistency of the restated
pytorch3d rasterisation rule (there is no pytorch3d here to pin it against: parity unpinned, see the oracle's header)."""
import numpy as np
import torch

from invertavatar_b200 import synth
from oracle import faceverse as o_fv


def test_rasterize_single_triangle_rule():
    """Pixel centres at NDC 1 - (2i+1)/S, +X left / +Y up (x_ndc = -x_world), strict inside test, nearest view z wins."""
    verts = np.array([[-0.5, -0.5, 0.0], [0.5, -0.5, 0.0], [-0.5, 0.5, 0.0],       # near triangle (z = 0)
                      [-0.9, -0.9, 1.0], [0.9, -0.9, 1.0], [-0.9, 0.9, 1.0]], dtype=np.float32)
    tri = np.array([[0, 1, 2], [3, 4, 5]])
    p2f, bary = o_fv.rasterize(verts, tri, size=8)
    # world (-1,-1) is pixel (row 0, col 0): x_ndc = +1 there
    assert p2f[0, 0] == 1 and p2f[7, 7] == -1
    assert p2f[2, 2] == 0 and p2f[1, 1] == 1                 # the near triangle hides the far one where both cover the pixel
    inside = p2f >= 0
    assert np.allclose(bary[inside].sum(-1), 1.0, atol=1e-6) and (bary[inside] > 0).all()


def test_make_driven_rendering_shapes_and_clamp():
    model, face_mask, trans_init = synth.faceverse_model(n=24, n_eye=6)
    co = synth.faceverse_coeffs(1)
    idc, exc, _, _, _, _, eye, _ = o_fv.split_coeffs(co, synth.FV_ID_DIMS, synth.FV_EXP_DIMS, synth.FV_TEX_DIMS)
    assert abs(float(o_fv.clamp_expression(exc)[0, -4]) - 0.6) < 1e-6 and eye.shape == (1, 4)
    attr = np.concatenate([model['uv_per_ver'] * 2 - 1, np.maximum(face_mask, 0)[:, None]], -1)
    img, p2f, verts = o_fv.make_driven_rendering(model, idc, exc, eye, trans_init, attr)
    assert tuple(img.shape) == (1, 256, 256, 3) and set(np.unique(img[..., 2].numpy())) <= {0.0, 1.0}
    assert float(img[..., :2].abs().max()) <= 1.0 and 0.1 < float((p2f >= 0).float().mean()) < 0.6
    # an eye rotation moves eye-ball vertices only
    eye2 = eye.clone(); eye2[0, 1] += 0.2
    v2 = o_fv.vertices(model, idc, o_fv.clamp_expression(exc), eye2, trans_init)
    moved = (v2 - verts).abs().amax(dim=1) > 1e-7
    vi = [int(v) for v in model['ver_inds']]
    assert bool(moved[vi[0]:vi[1]].any()) and not bool(moved[:vi[0]].any()) and not bool(moved[vi[1]:vi[2]].any())
