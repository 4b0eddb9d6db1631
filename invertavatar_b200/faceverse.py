"""Mesh-condition producer on the B200 engine (reference data_preprocess/FaceVerse/renderer.py ``Faceverse_manager`` and
the parts of data_preprocess/FaceVerse/FaceVerseModel_v3.py it uses).

``make_driven_rendering(drive_coeff)`` turns the 3DMM coefficients of a driving frame into the ``uvcoords_image`` [B,256,256,3]
that ``TriPlaneGenerator.synthesis`` takes as ``mesh_condition`` -- the step that runs once per frame immediately in front of
the generator (reenact_avatar_next3d.py:77-83 loads the same images from disk when they were pre-rendered; eval_seq.py:203
renders them online).  The reference does it with torch einsums and the pytorch3d mesh rasteriser; here it is three kernels of
libinvertavatar_b200.so (csrc/ia_mesh.cu): coefficient preparation (clamp / retarget / eye rotations), blend shapes + eye-ball
rotation + rigid-orthographic transform, and an orthographic z-buffer rasteriser with the resolve / crop / mask pass.

The FaceVerse v3 asset (``faceverse_v3_1.npy``) is not distributed with the reference repository; ``Faceverse_manager`` takes
the model either from that file (same location as the reference) or as an in-memory dict with the same keys."""
import os

import numpy as np
import torch

from . import _C
from . import runtime as rt

import ctypes as C


class FaceVerseModel:
    """The slice of FaceVerseModel_v3.FaceVerseModel the producer needs: bases with the loader's conventions applied
    (FaceVerseModel_v3.py:41-57: y/z flipped, scaled by 0.1, y lifted by 1), coefficient layout, vertex ranges."""

    def __init__(self, model_dict, device='cuda:0'):
        self.device = torch.device(device)
        nid, nexp = model_dict['idBase'].shape[-1], model_dict['exBase'].shape[-1]
        mean = torch.as_tensor(np.asarray(model_dict['meanshape']), dtype=torch.float32).reshape(-1, 3).clone()
        mean[:, [1, 2]] *= -1
        mean = mean * 0.1
        mean[:, 1] += 1
        idb = torch.as_tensor(np.asarray(model_dict['idBase']), dtype=torch.float32).reshape(-1, 3, nid).clone()
        idb[:, [1, 2]] *= -1
        exb = torch.as_tensor(np.asarray(model_dict['exBase']), dtype=torch.float32).reshape(-1, 3, nexp).clone()
        exb[:, [1, 2]] *= -1
        self.num_vertex = mean.shape[0]
        self.id_dims, self.exp_dims = nid, nexp
        self.tex_dims = int(model_dict['texBase'].shape[-1]) if 'texBase' in model_dict else 251
        self.all_dims = self.id_dims + self.tex_dims + self.exp_dims
        self.meanshape = mean.reshape(-1).to(self.device)                                   # [NV*3]
        self.idBase = (idb * 0.1).reshape(-1, nid).contiguous().to(self.device)             # [NV*3, id]   (rows = outputs of a GEMV)
        self.expBase_t = (exb * 0.1).reshape(-1, nexp).t().contiguous().to(self.device)     # [exp, NV*3]  (coalesced over vertices)
        self.tri = torch.as_tensor(np.asarray(model_dict['tri']), dtype=torch.int64).to(self.device)
        self.tri32 = self.tri.to(torch.int32).contiguous()
        self.ver_inds = [int(v) for v in model_dict['ver_inds']]

    def split_coeffs(self, coeffs):
        """FaceVerseModel_v3.py:139-153."""
        a = self.all_dims
        id_coeff = coeffs[:, :self.id_dims]
        exp_coeff = coeffs[:, self.id_dims:self.id_dims + self.exp_dims]
        tex_coeff = coeffs[:, self.id_dims + self.exp_dims:a]
        angles, gamma, translation = coeffs[:, a:a + 3], coeffs[:, a + 3:a + 30], coeffs[:, a + 30:a + 33]
        if coeffs.shape[1] == a + 36:
            eye_coeff, scale = coeffs[:, a + 33:], torch.ones_like(coeffs[:, -1:])
        else:
            eye_coeff, scale = coeffs[:, a + 33:-1], coeffs[:, -1:]
        return id_coeff, exp_coeff, tex_coeff, angles, gamma, translation, eye_coeff, scale

    def neutral_shape(self, id_coeff):
        """meanshape + idBase @ id -> [NV,3] (one GEMV of the library; per identity, not per frame)."""
        y = rt.fully_connected(id_coeff.reshape(1, -1).to(self.device).float(), self.idBase, self.meanshape)
        return y.reshape(self.num_vertex, 3)


class Faceverse_manager(object):
    """data_preprocess/FaceVerse/renderer.py:11-84.  ``model_dict`` / ``face_mask`` / ``trans_init`` default to the files the
    reference reads (data_preprocess/FaceVerse/v3/{faceverse_v3_1.npy, v31_face_mask_new.npy, fv2fl_30.npy})."""

    def __init__(self, device, base_coeff, model_dict=None, face_mask=None, trans_init=None, face_model_dir='data_preprocess/FaceVerse/v3'):
        self.device = torch.device(device)
        if model_dict is None:
            path = os.path.join(face_model_dir, 'faceverse_v3_1.npy')
            if not os.path.exists(path):
                raise FileNotFoundError(f'{path}: the FaceVerse v3 model is not distributed with the reference repository; '
                                        'pass model_dict=... or place the file there')
            model_dict = np.load(path, allow_pickle=True).item()
        if face_mask is None:
            face_mask = np.load(os.path.join(face_model_dir, 'v31_face_mask_new.npy'))
        if trans_init is None:
            trans_init = np.load(os.path.join(face_model_dir, 'fv2fl_30.npy'))
        self.render_res = 512
        self.orth_scale, self.orth_shift = 5.00, np.asarray([0, 0.005, 0.], dtype=np.float32)
        self.recon_model = FaceVerseModel(model_dict, device=self.device)
        vi = self.recon_model.ver_inds
        uv = np.asarray(model_dict['uv_per_ver'], dtype=np.float32).copy()
        idx = (uv[:, 1] > 0.273) * (uv[:, 1] < 0.727) * (uv[:, 0] > 0.195) * (uv[:, 0] < 0.805)      # enlarge the face region of the UV map (:24-26)
        uv[idx] = (uv[idx] - 0.5) * 1.4 + 0.5
        mask = np.asarray(face_mask, dtype=np.float32).reshape(-1).copy()
        mask[vi[0]:vi[2]] = 1                                                                       # eye-balls belong to the face (:29)
        self.vert_attr = torch.from_numpy(np.concatenate([uv * 2 - 1, mask[:, None]], axis=-1).astype(np.float32)).contiguous().to(self.device)
        T = np.asarray(trans_init, dtype=np.float32)
        # vert = vs @ T[:3,:3].T + T[:3,3]; (vert @ I + shift) * scale; batch_orth_proj with cam [1,0,0]; z *= -1   (:59-66)
        M = np.concatenate([T[:3, :3], (T[:3, 3] + self.orth_shift)[:, None]], axis=1) * self.orth_scale
        M[2] *= -1
        self.M = M.astype(np.float32)
        self.crop_param = [128, 114, 256, 256]            # left, top, width, height
        self.id_coeff = self.base_avatar_exp_coeff = None
        self._neutral = self._centres = None
        if base_coeff is not None:
            assert isinstance(base_coeff, torch.Tensor) and base_coeff.ndim == 1
            self.id_coeff, self.base_avatar_exp_coeff = self.recon_model.split_coeffs(base_coeff.to(self.device).unsqueeze(0))[:2]

    def _identity(self):
        """Per-identity state: neutral shape and eye-ball centres (recomputed when ``id_coeff`` is reassigned, eval_seq.py:192)."""
        key = (self.id_coeff.data_ptr(), self.id_coeff._version)
        if self._neutral is None or self._neutral[0] != key:
            neutral = self.recon_model.neutral_shape(self.id_coeff).contiguous()
            centres = torch.empty(6, dtype=torch.float32, device=self.device)
            vi = self.recon_model.ver_inds
            st = torch.cuda.current_stream(self.device).cuda_stream
            _C.check(_C.lib().ia_mesh_eye_centres(neutral.data_ptr(), vi[0], vi[1], vi[2], centres.data_ptr(), st), 'ia_mesh_eye_centres')
            self._neutral = (key, neutral, centres)
        return self._neutral[1], self._neutral[2]

    def driven_vertices(self, drive_coeff, base_drive_coeff=None):
        """[B, D] driving coefficients -> [B, NV, 3] vertices in the rasteriser's space (renderer.py:45-66)."""
        assert drive_coeff.ndim == 2 and self.id_coeff is not None
        m = self.recon_model
        coeff = drive_coeff.to(self.device).float().contiguous()
        B = coeff.shape[0]
        with torch.cuda.device(self.device):
            neutral, centres = self._identity()
            st = torch.cuda.current_stream(self.device).cuda_stream
            exp = torch.empty((B, m.exp_dims), dtype=torch.float32, device=self.device)
            eye_rot = torch.empty((B, 2, 9), dtype=torch.float32, device=self.device)
            bd = ba = None
            if base_drive_coeff is not None:
                bd = m.split_coeffs(base_drive_coeff.to(self.device).float())[1].reshape(-1).contiguous()
                ba = self.base_avatar_exp_coeff.reshape(-1).float().contiguous()
            _C.check(_C.lib().ia_mesh_coeffs(coeff.data_ptr(), coeff.stride(0), B, m.id_dims, m.exp_dims, m.all_dims + 33,
                                             None if bd is None else bd.data_ptr(), None if ba is None else ba.data_ptr(),
                                             exp.data_ptr(), eye_rot.data_ptr(), st), 'ia_mesh_coeffs')
            verts = torch.empty((B, m.num_vertex, 3), dtype=torch.float32, device=self.device)
            p = _C.BlendshapeParams()
            p.neutral, p.exp_basis_t, p.exp = neutral.data_ptr(), m.expBase_t.data_ptr(), exp.data_ptr()
            p.B, p.NV, p.exp_dims = B, m.num_vertex, m.exp_dims
            p.eye0, p.eye1, p.eye2 = m.ver_inds[0], m.ver_inds[1], m.ver_inds[2]
            p.eye_rot, p.eye_centre, p.verts = eye_rot.data_ptr(), centres.data_ptr(), verts.data_ptr()
            for i, v in enumerate(self.M.reshape(-1)):
                p.M[i] = float(v)
            _C.check(_C.lib().ia_blendshape(C.byref(p), st), 'ia_blendshape')
        return verts

    def rasterize_vertices(self, verts, return_pix_to_face=False):
        """[B, NV, 3] -> uvcoords_image [B, 256, 256, 3] (renderer.py:68-84 without the optional resize)."""
        m = self.recon_model
        B = verts.shape[0]
        left, top, w, h = self.crop_param
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            zbuf = torch.empty(int(_C.lib().ia_ortho_raster_scratch_bytes(B, self.render_res)) // 8, dtype=torch.int64, device=self.device)
            out = torch.empty((B, h, w, 3), dtype=torch.float32, device=self.device)
            p2f = torch.empty((B, h, w), dtype=torch.int32, device=self.device) if return_pix_to_face else None
            verts = verts.contiguous()
            p = _C.OrthoRasterParams(verts.data_ptr(), B, m.num_vertex, m.tri32.data_ptr(), int(m.tri32.shape[0]), self.vert_attr.data_ptr(),
                                     self.render_res, 10.0, 1e-6, left, top, w, h, zbuf.data_ptr(), out.data_ptr(),
                                     None if p2f is None else p2f.data_ptr())
            _C.check(_C.lib().ia_ortho_raster(C.byref(p), st), 'ia_ortho_raster')
        return (out, p2f) if return_pix_to_face else out

    def make_driven_rendering(self, drive_coeff, base_drive_coeff=None, res=None):
        """renderer.py:45-84.  ``res`` other than the 256-pixel crop size would need the reference's bilinear resize of the
        rendering; the inference scripts ask for 256 (eval_seq.py:203)."""
        if not (res is None or res == self.crop_param[2]):
            raise NotImplementedError('make_driven_rendering: only res=None / 256 (the size of the crop) is implemented')
        return self.rasterize_vertices(self.driven_vertices(drive_coeff, base_drive_coeff))
