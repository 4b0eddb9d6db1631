"""Synthetic, seeded inputs for the generator-forward hot path.

These are the inputs SURVEY.md section 8(d) specifies: latents from
``RandomState(i)`` (same rule as reference ``reenact_avatar_next3d.py:172``),
look-at cameras at a fixed radius (reference
``training_avatar_texture/camera_utils.py:58-87,105-148``), a UV mesh condition
with a mouth hole (so ``fill_mouth`` has something to fill), and the explicit
depth-jitter tensor that replaces ``torch.rand_like`` (reference
``volumetric_rendering/renderer.py:406``).

Everything here is plain CPU torch/numpy; tests, bench.py and the golden
generator all draw their inputs from this one place.
"""
import math

import numpy as np
import torch

FOV_DEG = 18.837
CAM_RADIUS = 2.7
CAM_PIVOT = (0.0, 0.0, 0.2)


def rendering_kwargs(depth_resolution=48, depth_resolution_importance=48):
    """The ``rendering_kwargs`` dict that travels inside reference pickles
    (``train_avatar_texture.py:320-348``)."""
    return {
        'image_resolution': 512,
        'disparity_space_sampling': False,
        'clamp_mode': 'softplus',
        'superresolution_module': 'training_avatar_texture.superresolution.SuperresolutionHybrid8XDC',
        'c_gen_conditioning_zero': False,
        'gpc_reg_prob': None,
        'c_scale': 1,
        'superresolution_noise_mode': 'none',
        'density_reg': 0.25,
        'density_reg_p_dist': 0.004,
        'reg_type': 'l1',
        'decoder_lr_mul': 1,
        'sr_antialias': True,
        'depth_resolution': depth_resolution,
        'depth_resolution_importance': depth_resolution_importance,
        'ray_start': 2.25,
        'ray_end': 3.3,
        'box_warp': 1,
        'avg_camera_radius': CAM_RADIUS,
        'avg_camera_pivot': list(CAM_PIVOT),
    }


def generator_kwargs(depth_resolution=48, depth_resolution_importance=48):
    """Constructor kwargs of the Next3D++ generator at the ``train_avatar_texture.py``
    defaults (``:256,274-276,304,355,365-367,392-393``)."""
    return dict(
        z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3,
        sr_num_fp16_res=4,
        mapping_kwargs={'num_layers': 2},
        rendering_kwargs=rendering_kwargs(depth_resolution, depth_resolution_importance),
        sr_kwargs={'channel_base': 32768, 'channel_max': 512, 'fused_modconv_default': 'inference_only'},
        channel_base=32768, channel_max=512, fused_modconv_default='inference_only',
        num_fp16_res=0, conv_clamp=None,
    )


def _normalize(v):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True)


def lookat_cam2world(h, v, pivot=CAM_PIVOT, radius=CAM_RADIUS):
    """Look-at pose, y-up, no roll (semantics of ``LookAtPoseSampler.sample`` +
    ``create_cam2world_matrix``, ``camera_utils.py:58-87,119-137``)."""
    h = torch.as_tensor(h, dtype=torch.float32).reshape(-1, 1)
    v = torch.as_tensor(v, dtype=torch.float32).reshape(-1, 1)
    v = torch.clamp(v, 1e-5, math.pi - 1e-5)
    phi = torch.arccos(1 - 2 * (v / math.pi))
    origin = torch.zeros(h.shape[0], 3)
    origin[:, 0:1] = radius * torch.sin(phi) * torch.cos(math.pi - h)
    origin[:, 2:3] = radius * torch.sin(phi) * torch.sin(math.pi - h)
    origin[:, 1:2] = radius * torch.cos(phi)
    fwd = _normalize(torch.tensor(pivot, dtype=torch.float32) - origin)
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    right = -_normalize(torch.cross(up, fwd, dim=-1))
    up = _normalize(torch.cross(fwd, right, dim=-1))
    rot = torch.eye(4).repeat(h.shape[0], 1, 1)
    rot[:, :3, :3] = torch.stack((right, up, fwd), dim=-1)
    trans = torch.eye(4).repeat(h.shape[0], 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


def fov_to_intrinsics(fov_degrees=FOV_DEG):
    """Normalised pinhole intrinsics (``camera_utils.py:140-148``; note the
    reference's 3.14159 / 1.414 constants)."""
    focal = float(1 / (math.tan(fov_degrees * 3.14159 / 360) * 1.414))
    return torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=torch.float32)


def cameras(batch, first=0):
    """[B,25] labels: c2w(16) | K(9). yaw,pitch ~ U(-0.4,0.4),U(-0.2,0.2) from RandomState(1000+i)."""
    hs, vs = [], []
    for i in range(first, first + batch):
        rs = np.random.RandomState(1000 + i)
        hs.append(math.pi / 2 + rs.uniform(-0.4, 0.4))
        vs.append(math.pi / 2 + rs.uniform(-0.2, 0.2))
    c2w = lookat_cam2world(hs, vs)
    K = fov_to_intrinsics().reshape(1, 9).repeat(batch, 1)
    return torch.cat([c2w.reshape(batch, 16), K], dim=1)


def frontal_camera(batch=1):
    c2w = lookat_cam2world([math.pi / 2] * batch, [math.pi / 2] * batch)
    K = fov_to_intrinsics().reshape(1, 9).repeat(batch, 1)
    return torch.cat([c2w.reshape(batch, 16), K], dim=1)


def latents(batch, first=0, z_dim=512):
    return torch.from_numpy(np.concatenate(
        [np.random.RandomState(i).randn(1, z_dim) for i in range(first, first + batch)])).float()


def uvcoords_image(batch, first=0, res=256):
    """[B,res,res,3]: (u,v) in [-1,1] = smooth warp of a normalised grid; mask = face
    ellipse minus an enclosed mouth ellipse, binarised at 0.5
    (``reenact_avatar_next3d.py:79``)."""
    out = []
    lin = (torch.arange(res, dtype=torch.float32) + 0.5) / res * 2 - 1
    yy, xx = torch.meshgrid(lin, lin, indexing='ij')
    for i in range(first, first + batch):
        rs = np.random.RandomState(2000 + i)
        ph = rs.uniform(0, 2 * math.pi, size=4).astype(np.float32)
        u = xx + 0.1 * torch.sin(3.0 * yy + float(ph[0])) * torch.cos(2.0 * xx + float(ph[1]))
        v = yy + 0.1 * torch.sin(2.5 * xx + float(ph[2])) * torch.cos(3.5 * yy + float(ph[3]))
        face = ((xx / 0.55) ** 2 + (yy / 0.7) ** 2) < 1
        mx, my = float(rs.uniform(-0.03, 0.03)), 0.3 + float(rs.uniform(-0.03, 0.03))
        mouth = (((xx - mx) / 0.12) ** 2 + ((yy - my) / 0.05) ** 2) < 1
        mask = (face & ~mouth).float()
        out.append(torch.stack([u.clamp(-1, 1), v.clamp(-1, 1), mask], dim=-1))
    return torch.stack(out)


def depth_jitter(batch, rays, depth_resolution, seed=7):
    """U[0,1) tensor [B,rays,D,1] that both the oracle and the CUDA renderer consume in
    place of ``torch.rand_like`` (``renderer.py:406``). One generator per sample so the
    values of sample i do not depend on the batch size (shard invariance)."""
    out = []
    for i in range(batch):
        g = torch.Generator(device='cpu').manual_seed(seed * 100003 + i)
        out.append(torch.rand(rays, depth_resolution, 1, generator=g))
    return torch.stack(out)


def importance_u(batch, rays, n_importance, seed=11):
    """U[0,1) tensor [B*rays, n_importance] replacing ``torch.rand`` when evaluation=False
    (``renderer.py:453``)."""
    out = []
    for i in range(batch):
        g = torch.Generator(device='cpu').manual_seed(seed * 100003 + i)
        out.append(torch.rand(rays, n_importance, generator=g))
    return torch.cat(out)


def randomize_noise_and_wavg(module_or_sd, seed=123):
    """Random init leaves noise_strength=0 and w_avg=0; give them values so the noise and
    truncation paths are exercised. Accepts a module or a state-dict."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    with torch.no_grad():
        for k in sorted(sd.keys()):
            v = sd[k]
            if k.endswith('noise_strength'):
                v.copy_((torch.rand([], generator=g) * 0.1).to(v.device))
            elif k.endswith('w_avg'):
                v.copy_(torch.randn(v.shape, generator=g).to(v.device))
            elif k.endswith('.bias') and ('conv' in k or 'torgb' in k) and 'affine' not in k:
                v.copy_((torch.randn(v.shape, generator=g) * 0.1).to(v.device))
    return module_or_sd


def randomize_encoder(net, seed=321):
    """Random init leaves every BatchNorm at identity (weight 1, bias 0, running mean 0 / var 1) and every PReLU at
    0.25; give them values so that the normalisation and activation paths of the inversion encoder are exercised.
    Works on the reference modules and on this package's (both keep torch.nn.BatchNorm2d / PReLU parameter holders)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_((0.8 + 0.4 * torch.rand(m.weight.shape, generator=g)).to(m.weight.device))
                m.bias.copy_((0.1 * torch.randn(m.bias.shape, generator=g)).to(m.bias.device))
                m.running_mean.copy_((0.1 * torch.randn(m.running_mean.shape, generator=g)).to(m.running_mean.device))
                m.running_var.copy_((0.5 + torch.rand(m.running_var.shape, generator=g)).to(m.running_var.device))
            elif isinstance(m, torch.nn.PReLU):
                m.weight.copy_((0.1 + 0.3 * torch.rand(m.weight.shape, generator=g)).to(m.weight.device))
    return net


def randomize_by_name(net, seed=77):
    """Overwrite EVERY parameter and floating-point buffer of ``net`` with values drawn from a generator seeded by the tensor's
    own state-dict name (crc32(name) ^ seed), so that two implementations with the same state-dict names and shapes end up
    with identical weights whatever order their constructors draw random numbers in (used for the SegFormer-style one-shot
    encoder, SURVEY 8f-4, whose reference constructors re-initialise nested modules several times).  Magnitudes keep
    activations O(1): matrices / filters N(0, 1/fan_in), biases 0.1 N(0,1), norm weights in [0.8, 1.2], PReLU slopes in
    [0.1, 0.4], running variances in [0.5, 1.5]."""
    import zlib
    kinds = {}
    for mname, m in net.named_modules():
        for pname, _ in list(m.named_parameters(recurse=False)) + list(m.named_buffers(recurse=False)):
            kinds[(mname + '.' if mname else '') + pname] = type(m).__name__
    with torch.no_grad():
        for name, v in net.state_dict().items():
            if not v.dtype.is_floating_point:
                continue
            g = torch.Generator(device='cpu').manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
            leaf = name.rsplit('.', 1)[-1]
            kind = kinds.get(name, '')
            if v.ndim >= 2:
                fan_in = v[0].numel()
                val = torch.randn(v.shape, generator=g) * (1.0 / fan_in) ** 0.5
            elif leaf == 'running_var':
                val = 0.5 + torch.rand(v.shape, generator=g)
            elif leaf == 'weight' and kind == 'PReLU':
                val = 0.1 + 0.3 * torch.rand(v.shape, generator=g)
            elif leaf == 'weight':
                val = 0.8 + 0.4 * torch.rand(v.shape, generator=g)
            else:
                val = 0.1 * torch.randn(v.shape, generator=g)
            v.copy_(val.to(v.device))
    return net


def encoder_inputs(T, seed=31):
    """SURVEY 8(d) encoder config: images clamp(N(0,0.5),-1,1) [T,3,512,512]; uv [T,6,256,256] = (random texture 3 ch,
    the UV mesh condition image 3 ch); cameras and mesh conditions of frames 0..T-1."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    image = (0.5 * torch.randn(T, 3, 512, 512, generator=g)).clamp(-1, 1)
    tex = (0.5 * torch.randn(T, 3, 256, 256, generator=g)).clamp(-1, 1)
    uvimg = uvcoords_image(T)
    uv = torch.cat([tex, uvimg.permute(0, 3, 1, 2)], dim=1)
    return {'image': image, 'uv': uv}, cameras(T), {'uvcoords_image': uvimg}


# ---- synthetic FaceVerse-style 3DMM (mesh-condition producer) -------------------------------------------
# The real asset (data_preprocess/FaceVerse/v3/faceverse_v3_1.npy) is not part of the reference repository; this generator
# builds a model dict with the same keys, conventions and dimensions (150 identity + 171 expression blend shapes, eye-ball
# vertex ranges in ``ver_inds``, per-vertex UVs) over a small dome-shaped mesh, so that the producer can be exercised end to end.
FV_ID_DIMS, FV_EXP_DIMS, FV_TEX_DIMS = 150, 171, 251


def faceverse_model(n=64, n_eye=12, seed=5):
    """-> (model dict as np.load('faceverse_v3_1.npy').item() would give, face_mask [NV], trans_init [4,4])."""
    rs = np.random.RandomState(seed)

    def grid(m, cx, cy, sx, sy, z0, depth):
        a, b = np.meshgrid(np.linspace(-1, 1, m), np.linspace(-1, 1, m))
        x, y = cx + sx * a, cy + sy * b
        z = z0 - depth * (1 - 0.5 * (a * a + b * b))
        idx = np.arange(m * m).reshape(m, m)
        t0 = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:]], -1).reshape(-1, 3)
        t1 = np.stack([idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 3)
        return np.stack([x, y, z], -1).reshape(-1, 3), np.concatenate([t0, t1]), np.stack([a, b], -1).reshape(-1, 2)
    face_v, face_t, face_ab = grid(n, 0.0, 0.0, 0.9, 1.0, 0.0, 0.6)
    le_v, le_t, le_ab = grid(n_eye, -0.35, -0.3, 0.12, 0.08, -0.50, 0.06)
    re_v, re_t, re_ab = grid(n_eye, 0.35, -0.3, 0.12, 0.08, -0.50, 0.06)
    nf, ne = face_v.shape[0], le_v.shape[0]
    verts = np.concatenate([face_v, le_v, re_v]).astype(np.float32)
    tri = np.concatenate([face_t, le_t + nf, re_t + nf + ne]).astype(np.int64)
    ab = np.concatenate([face_ab, 0.12 * le_ab + [-0.35, -0.3], 0.12 * re_ab + [0.35, -0.3]]).astype(np.float32)
    nv = verts.shape[0]

    def basis(dims, amp):
        out = np.zeros((nv, 3, dims), dtype=np.float32)
        for j in range(dims):
            f = rs.uniform(0.5, 3.5, size=2)
            ph = rs.uniform(0, 2 * math.pi)
            d = rs.randn(3).astype(np.float32)
            d /= np.linalg.norm(d)
            out[:, :, j] = (amp * np.sin(f[0] * ab[:, 0] + f[1] * ab[:, 1] + ph))[:, None] * d[None, :]
        return out.reshape(nv * 3, dims)
    model = {
        # raw model space: the loader flips y/z, scales by 0.1 and lifts y by 1 (FaceVerseModel_v3.py:41-57); stored pre-inverted
        'meanshape': (verts * np.array([1, -1, -1], dtype=np.float32)).reshape(-1),
        'idBase': basis(FV_ID_DIMS, 0.02), 'exBase': basis(FV_EXP_DIMS, 0.03),
        'tri': tri, 'ver_inds': np.array([nf, nf + ne, nf + 2 * ne], dtype=np.int64),
        'uv_per_ver': ((ab + 1) / 2).astype(np.float32),
    }
    face_mask = np.concatenate([(face_ab[:, 0] ** 2 + face_ab[:, 1] ** 2 < 0.9).astype(np.float32), np.zeros(2 * ne, dtype=np.float32)])
    trans_init = np.eye(4, dtype=np.float32)
    trans_init[1, 3] = -1.0                      # undo the loader's y lift: the head ends up centred in the orthographic window
    return model, face_mask, trans_init


def faceverse_coeffs(batch, first=0, with_scale=True):
    """Driving coefficient vectors [B, 150 + 171 + 251 + 3 + 27 + 3 + 4 (+1)] in FaceVerseModel.split_coeffs order
    (FaceVerseModel_v3.py:139-153): small random identity / expression, eye rotations within +-0.3 rad."""
    out = []
    for i in range(first, first + batch):
        rs = np.random.RandomState(3000 + i)
        c = np.concatenate([0.5 * rs.randn(FV_ID_DIMS), 0.5 * rs.randn(FV_EXP_DIMS), np.zeros(FV_TEX_DIMS), np.zeros(3), np.zeros(27), np.zeros(3),
                            rs.uniform(-0.3, 0.3, size=4), np.ones(1 if with_scale else 0)])
        c[FV_ID_DIMS + FV_EXP_DIMS - 4] = 2.0          # outside the clamp range of renderer.py:48
        out.append(c)
    return torch.from_numpy(np.stack(out)).float()
