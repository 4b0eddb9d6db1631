"""Oracle restatement of the mesh-condition producer.  TEST INFRASTRUCTURE ONLY.

Follows reference data_preprocess/FaceVerse/renderer.py:45-84 (``Faceverse_manager.make_driven_rendering``),
data_preprocess/FaceVerse/FaceVerseModel_v3.py:139-153 (``split_coeffs``), :237-244 (``get_vs``), :252-264 (eye centres),
:303-325 (``compute_eye_rotation_matrix``) and training_avatar_texture/volumetric_rendering/renderer.py:556-571
(``render_after_rasterize``), :636-646 (``batch_orth_proj``).

PARITY UNPINNED for the rasterisation step: the reference rasterises with pytorch3d (``MeshRasterizer`` of an
``OrthographicCameras(R=I, T=(0,0,10), focal_length=(-1,-1), principal_point=(0,0), in_ndc=True)``, 512^2,
``blur_radius=1e-6``, ``faces_per_pixel=1``; ortho_renderer.py:52-100), a dependency that is neither vendored nor pinned
(absent from environment.yml) nor installed here.  ``rasterize`` below restates pytorch3d's published naive algorithm
(pytorch3d/csrc/rasterize_meshes/rasterize_meshes.cu ``RasterizeMeshesNaiveCudaKernel`` / ``CheckPixelInsideFace`` and
csrc/utils/geometry_utils.cuh: pixel centres at NDC ``1 - (2i+1)/S`` with +X left / +Y up, strict-positive barycentric inside
test, the nearest positive view-space z wins, ties to the lower face index) -- it has not been run against pytorch3d.
The blend-shape / transform part is plain tensor algebra and is exact."""
import math

import numpy as np
import torch


def split_coeffs(coeffs, id_dims, exp_dims, tex_dims):
    """FaceVerseModel_v3.py:139-153."""
    all_dims = id_dims + exp_dims + tex_dims
    id_c = coeffs[:, :id_dims]
    exp_c = coeffs[:, id_dims:id_dims + exp_dims]
    tex_c = coeffs[:, id_dims + exp_dims:all_dims]
    angles = coeffs[:, all_dims:all_dims + 3]
    gamma = coeffs[:, all_dims + 3:all_dims + 30]
    trans = coeffs[:, all_dims + 30:all_dims + 33]
    if coeffs.shape[1] == all_dims + 36:
        eye = coeffs[:, all_dims + 33:]
        scale = torch.ones_like(coeffs[:, -1:])
    else:
        eye = coeffs[:, all_dims + 33:-1]
        scale = coeffs[:, -1:]
    return id_c, exp_c, tex_c, angles, gamma, trans, eye, scale


def preprocess_model(model):
    """FaceVerseModel_v3.py:41-57: flip y/z, scale 0.1, lift y by 1."""
    mean = torch.as_tensor(model['meanshape'], dtype=torch.float32).reshape(-1, 3).clone()
    mean[:, [1, 2]] *= -1
    mean = mean * 0.1
    mean[:, 1] += 1
    nid, nexp = model['idBase'].shape[-1], model['exBase'].shape[-1]
    idb = torch.as_tensor(model['idBase'], dtype=torch.float32).reshape(-1, 3, nid).clone()
    idb[:, [1, 2]] *= -1
    exb = torch.as_tensor(model['exBase'], dtype=torch.float32).reshape(-1, 3, nexp).clone()
    exb[:, [1, 2]] *= -1
    return mean.reshape(1, -1), (idb * 0.1).reshape(-1, nid), (exb * 0.1).reshape(-1, nexp)


def eye_rotation(eye2):
    """FaceVerseModel_v3.py:303-325: R = Ry(eye[1]) @ Rx(eye[0]) (batch 1)."""
    sx, sy, cx, cy = math.sin(float(eye2[0])), math.sin(float(eye2[1])), math.cos(float(eye2[0])), math.cos(float(eye2[1]))
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float32)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float32)
    return ry @ rx


def vertices(model, id_coeff, exp_coeff, eye_coeff, trans_init, orth_scale=5.0, orth_shift=(0.0, 0.005, 0.0)):
    """renderer.py:45-68 up to the rasteriser: blend shapes, eye-ball rotation about the identity's eye centres, the rigid
    fv2fl transform, orthographic scale/shift, z flip -> [NV,3] in the rasteriser's world space."""
    mean, idb, exb = preprocess_model(model)
    vi = [int(v) for v in model['ver_inds']]
    shape = (idb @ id_coeff.reshape(-1) + exb @ exp_coeff.reshape(-1) + mean.reshape(-1)).reshape(-1, 3)
    neutral = (idb @ id_coeff.reshape(-1) + mean.reshape(-1)).reshape(-1, 3)
    for k, (a, b) in enumerate(((vi[0], vi[1]), (vi[1], vi[2]))):
        centre = neutral[a:b].clone()
        centre[:, 2] += 0.005
        centre = centre.mean(dim=0, keepdim=True)
        R = eye_rotation(eye_coeff.reshape(-1)[2 * k:2 * k + 2])
        shape[a:b] = (shape[a:b] - centre) @ R + centre
    T = torch.as_tensor(trans_init, dtype=torch.float32)
    vert = shape @ T[:3, :3].T + T[:3, 3:].T
    out = (vert + torch.tensor(orth_shift, dtype=torch.float32)) * orth_scale       # tform = identity, cam = [1, 0, 0]
    out[:, 2] *= -1
    return out


def clamp_expression(exp_coeff):
    """renderer.py:48-49."""
    e = exp_coeff.clone()
    e[:, -4] = max(min(float(e[0, -4]), 0.6), -0.75)
    e[:, -2] = max(min(float(e[0, -2]), 0.75), -0.75)
    return e


def _edge(ax, ay, bx, by, px, py):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def rasterize(verts, tri, size=512, cam_z=10.0, blur_radius=1e-6):
    """pix_to_face [S,S] (int64, -1 = background) and barycentrics [S,S,3] of the nearest face per pixel (see module docstring).
    verts [NV,3] world; the camera looks down +z from z = -cam_z with focal (-1,-1): NDC x = -x_world, y = -y_world."""
    v = np.asarray(verts, dtype=np.float32)
    f = np.asarray(tri, dtype=np.int64)
    xn, yn, zv = -v[:, 0], -v[:, 1], v[:, 2] + np.float32(cam_z)
    S = size
    pix2face = -np.ones((S, S), dtype=np.int64)
    zbuf = np.full((S, S), np.inf, dtype=np.float32)
    bary = np.zeros((S, S, 3), dtype=np.float32)
    centre = (np.float32(1.0) - (np.float32(2.0) * np.arange(S, dtype=np.float32) + np.float32(1.0)) / np.float32(S))   # pixel i -> NDC
    rad = np.float32(math.sqrt(blur_radius))
    for fi in range(f.shape[0]):
        i0, i1, i2 = f[fi]
        x0, y0, x1, y1, x2, y2 = xn[i0], yn[i0], xn[i1], yn[i1], xn[i2], yn[i2]
        area = _edge(x0, y0, x1, y1, x2, y2)
        if abs(float(area)) <= 1e-8:
            continue
        xmin, xmax = min(x0, x1, x2) - rad, max(x0, x1, x2) + rad
        ymin, ymax = min(y0, y1, y2) - rad, max(y0, y1, y2) + rad
        cols = np.nonzero((centre >= xmin) & (centre <= xmax))[0]
        rows = np.nonzero((centre >= ymin) & (centre <= ymax))[0]
        if cols.size == 0 or rows.size == 0:
            continue
        px, py = np.meshgrid(centre[cols], centre[rows])
        w0 = _edge(x1, y1, x2, y2, px, py) / area
        w1 = _edge(x2, y2, x0, y0, px, py) / area
        w2 = _edge(x0, y0, x1, y1, px, py) / area
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        if not inside.any():
            continue
        pz = w0 * zv[i0] + w1 * zv[i1] + w2 * zv[i2]
        rr, cc = np.meshgrid(rows, cols, indexing='ij')
        take = inside & (pz >= 0) & (pz < zbuf[rr, cc])
        if take.any():
            r_, c_ = rr[take], cc[take]
            zbuf[r_, c_] = pz[take]
            pix2face[r_, c_] = fi
            bary[r_, c_, 0], bary[r_, c_, 1], bary[r_, c_, 2] = w0[take], w1[take], w2[take]
    return pix2face, bary


def make_driven_rendering(model, id_coeff, exp_coeff, eye_coeff, trans_init, vert_attr, size=512, crop=(128, 114, 256, 256)):
    """renderer.py:45-84: -> uvcoords_image [1,h,w,3] (u, v in [-1,1] zeroed outside the face mask, binarised mask).
    vert_attr [NV,3] = (u*2-1, v*2-1, face mask) per vertex (renderer.py:23-33)."""
    verts = vertices(model, id_coeff, clamp_expression(exp_coeff), eye_coeff, trans_init)
    tri = np.asarray(model['tri'], dtype=np.int64)
    p2f, bary = rasterize(verts.numpy(), tri, size=size)
    attr = np.asarray(vert_attr, dtype=np.float32)
    vis = (p2f > -1)
    fa = attr[tri[np.where(vis, p2f, 0)]]                                # [S,S,3 verts,3 attrs]
    vals = (bary[..., None] * fa).sum(axis=-2)
    vals[~vis] = 0
    rendering = np.concatenate([vals, vis[..., None].astype(np.float32)], axis=-1)     # u, v, mask, vis
    render_mask = rendering[..., 3:4] * rendering[..., 2:3]
    rendering = rendering * render_mask
    left, top, w, h = crop
    out = rendering[top:top + h, left:left + w, :3].copy()
    out[..., 2] = (out[..., 2] >= 0.5).astype(np.float32)
    return torch.from_numpy(out).unsqueeze(0), torch.from_numpy(p2f), verts
