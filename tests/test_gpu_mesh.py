"""GPU parity of the mesh-condition producer (Faceverse_manager.make_driven_rendering) against the oracle
restatement (oracle/faceverse.py) on the synthetic FaceVerse-style model of invertavatar_b200.synth (the real asset is not
part of the reference repository; the rasterisation rule is pytorch3d's, which is not installed: parity with it is unpinned)."""
import numpy as np
import pytest
import torch

from invertavatar_b200 import synth
from invertavatar_b200.faceverse import Faceverse_manager
from oracle import faceverse as o_fv

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def setup():
    model, face_mask, trans_init = synth.faceverse_model()
    coeffs = synth.faceverse_coeffs(3)
    mgr = Faceverse_manager(DEV, coeffs[0], model_dict=model, face_mask=face_mask, trans_init=trans_init)
    return model, face_mask, trans_init, coeffs, mgr


def _oracle(model, face_mask, trans_init, base, drive, mgr):
    idc = o_fv.split_coeffs(base[None], synth.FV_ID_DIMS, synth.FV_EXP_DIMS, synth.FV_TEX_DIMS)[0]
    _, exc, _, _, _, _, eye, _ = o_fv.split_coeffs(drive[None], synth.FV_ID_DIMS, synth.FV_EXP_DIMS, synth.FV_TEX_DIMS)
    return o_fv.make_driven_rendering(model, idc, exc, eye, trans_init, mgr.vert_attr.cpu().numpy())


def test_driven_vertices_vs_oracle(setup):
    """Blend shapes (150 identity + 171 expression dims), expression clamp, eye-ball rotations, fv2fl / orthographic transform."""
    model, face_mask, trans_init, coeffs, mgr = setup
    got = mgr.driven_vertices(coeffs.to(DEV)).cpu()
    assert tuple(got.shape) == (3, mgr.recon_model.num_vertex, 3)
    for b in range(3):
        _, _, want = _oracle(model, face_mask, trans_init, coeffs[0], coeffs[b], mgr)
        assert float((got[b] - want).abs().max()) <= 2e-5
    # the clamp of renderer.py:48 is active in the synthetic coefficients
    assert float(coeffs[0, synth.FV_ID_DIMS + synth.FV_EXP_DIMS - 4]) > 0.6


def test_rasteriser_bit_exact_on_oracle_vertices(setup):
    """Given the same vertices the z-buffer rasteriser + resolve pass reproduces the float32 oracle bit for bit: face indices,
    barycentric interpolation of (u, v, mask), render-mask product, crop, binarised mask."""
    model, face_mask, trans_init, coeffs, mgr = setup
    for b in (0, 2):
        want_img, want_p2f, verts = _oracle(model, face_mask, trans_init, coeffs[0], coeffs[b], mgr)
        img, p2f = mgr.rasterize_vertices(verts[None].to(DEV), return_pix_to_face=True)
        left, top, w, h = mgr.crop_param
        assert torch.equal(p2f[0].cpu().long(), want_p2f[top:top + h, left:left + w])
        assert torch.equal(img.cpu(), want_img)
        assert 0.2 < float(img[..., 2].mean()) < 0.9            # the face covers a sensible part of the crop


def test_make_driven_rendering_end_to_end(setup):
    """Coefficients -> uvcoords_image through the public method; vertices differ from the oracle's in the last bits (different
    summation order), so a few pixels on triangle edges pick the neighbouring face and a few on the face-mask contour fall on
    the other side of the 0.5 threshold."""
    model, face_mask, trans_init, coeffs, mgr = setup
    got = mgr.make_driven_rendering(coeffs.to(DEV), res=256).cpu()
    assert tuple(got.shape) == (3, 256, 256, 3) and set(np.unique(got[..., 2].numpy())) <= {0.0, 1.0}
    for b in range(3):
        want, _, _ = _oracle(model, face_mask, trans_init, coeffs[0], coeffs[b], mgr)
        bad = ((got[b] - want[0]).abs().amax(dim=-1) > 1e-4).float().mean()
        assert float(bad) <= 1e-2, float(bad)
        assert float((got[b] - want[0]).abs().mean()) <= 1e-3
    # retargeting branch (renderer.py:50-53): driving with base_drive == drive reproduces the avatar's own expression
    same = mgr.make_driven_rendering(coeffs[1:2].to(DEV), base_drive_coeff=coeffs[1:2].to(DEV))
    own = mgr.make_driven_rendering(coeffs[0:1].to(DEV))
    eye_free = ((same - own).abs().amax(dim=-1) > 1e-4).float().mean()      # (eye rotations still come from the driving frame)
    assert float(eye_free) < 0.05
    with pytest.raises(NotImplementedError):
        mgr.make_driven_rendering(coeffs[:1].to(DEV), res=128)


def test_generator_consumes_the_produced_condition(setup):
    """The produced uvcoords_image drives TriPlaneGenerator.synthesis (the seam of eval_seq.py:203-212)."""
    import copy
    from common import build_generator
    model, face_mask, trans_init, coeffs, mgr = setup
    G = copy.deepcopy(build_generator(16, 16)).to(DEV)
    uv = mgr.make_driven_rendering(coeffs[:1].to(DEV), res=256)
    with torch.no_grad():
        ws = G.mapping(synth.latents(1).to(DEV), synth.frontal_camera(1).to(DEV), truncation_psi=0.7, truncation_cutoff=14)
        img = G.synthesis(ws, synth.cameras(1).to(DEV), {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True)['image']
    assert tuple(img.shape) == (1, 3, 512, 512) and bool(torch.isfinite(img).all())
