"""Volume-rendering front end (reference training_avatar_texture/volumetric_rendering/{renderer,ray_sampler,ray_marcher}.py).

The reference evaluates the renderer as ~60 tensor ops over [B, rays*samples, ...] intermediates; here the whole of
``ImportanceRenderer_bsMotion.forward`` is one persistent kernel (csrc/ia_render.cu).  The two random draws of the
reference -- coarse depth jitter (renderer.py:406, drawn even when evaluation=True) and the importance ``u`` when
evaluation=False (renderer.py:453) -- are explicit tensors: set ``renderer.depth_jitter`` / ``renderer.importance_u``
(consumed by the next forward) to reproduce a specific draw, otherwise they are sampled on the device."""
import math

import torch

from . import runtime as rt


def generate_planes(return_inv=True):
    """Tri-plane axis matrices, reference renderer.py:30-48 (projection is hard-wired in the kernel:
    plane 0 -> (x,y), plane 1 -> (x,z), plane 2 -> (z,x))."""
    planes = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                           [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                           [[0, 0, 1], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)
    return torch.linalg.inv(planes) if return_inv else planes


def fill_mouth(images, blur_mouth_edge=False):
    """[B,1,H,W] alpha -> (clip(alpha + mouth_mask, 0, 1), mouth_mask); GPU flood fill instead of the reference's
    host-side cv2.floodFill (renderer.py:716-741).  blur_mouth_edge (erode + blur of the returned mask) is not on the
    generator hot path (rasterize passes False, triplane_v20.py:323)."""
    if blur_mouth_edge:
        raise NotImplementedError('fill_mouth(blur_mouth_edge=True) is not used by the generator forward')
    full, mouth, _ = rt.fill_mouth(images)
    return full.unsqueeze(1), mouth.unsqueeze(1)


def sample_from_planes(plane_axes, plane_features, coordinates, mode='bilinear', padding_mode='zeros', box_warp=None, debug=False):
    """[N,3,C,H,W] planes sampled at [N,M,3] points -> [N,3,M,C] (renderer.py:85-97)."""
    assert padding_mode == 'zeros' and mode == 'bilinear'
    N, n_planes, Cc, H, W = plane_features.shape
    M = coordinates.shape[1]
    coords = (2 / box_warp) * coordinates
    outs = []
    pick = [(0, 1), (0, 2), (2, 0)]
    for p in range(n_planes):
        grid = torch.stack([coords[..., pick[p][0]], coords[..., pick[p][1]]], dim=-1).reshape(N, 1, M, 2)
        feat = rt.to_nhwc(plane_features[:, p])
        outs.append(rt.grid_sample_nhwc(feat, grid).reshape(N, M, Cc))
    return torch.stack(outs, dim=1)


class RaySampler_zxc(torch.nn.Module):
    """ray_sampler.py:65-107."""

    def __init__(self):
        super().__init__()

    def forward(self, cam2world_matrix, cam_K, resolution, normalize=True):
        if not normalize:
            raise NotImplementedError('RaySampler_zxc(normalize=False) is not used by the generator forward')
        B = cam2world_matrix.shape[0]
        cam = torch.cat([cam2world_matrix.reshape(B, 16), cam_K.reshape(B, 9)], dim=1)
        return rt.ray_sampler(cam, int(resolution))


RaySampler = RaySampler_zxc


class MipRayMarcher2(torch.nn.Module):
    """ray_marcher.py:19-62.  Inside the generator the marcher is fused into the render kernel; this module is the reference's
    standalone API (one kernel, ia_ray_march) for callers that march their own sorted samples."""

    def __init__(self):
        super().__init__()

    def run_forward(self, colors, densities, depths, rendering_options):
        assert rendering_options.get('clamp_mode', 'softplus') == 'softplus', 'MipRayMarcher only supports `clamp_mode`=`softplus`!'
        return rt.ray_march(colors, densities, depths, white_back=bool(rendering_options.get('white_back', False)))

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)


def _decoder_weights(decoder):
    net = decoder.net
    w1, b1, w2, b2 = net[0].weight, net[0].bias, net[2].weight, net[2].bias
    if tuple(w1.shape) != (64, 32) or tuple(w2.shape) != (33, 64):
        raise NotImplementedError(f'render kernel is specialised for the 32->64->33 OSG decoder, got {tuple(w1.shape)}, {tuple(w2.shape)}')
    lr1 = getattr(net[0], 'bias_gain', 1)
    if lr1 != 1 or getattr(net[2], 'bias_gain', 1) != 1:
        raise NotImplementedError('decoder_lr_mul != 1 is not supported by the fused renderer')
    return w1, b1, w2, b2


class ImportanceRenderer_bsMotion(torch.nn.Module):
    """renderer.py:295-469."""

    def __init__(self):
        super().__init__()
        self.ray_marcher = MipRayMarcher2()
        self.plane_axes = generate_planes()
        self.depth_jitter = None     # [B, rays, Dc(,1)] U[0,1) (or a list: one per forward); consumed by the next forward
        self.importance_u = None     # [B*rays, Df] U[0,1) (or a list); consumed by the next forward when evaluation=False
        self.fixed_jitter = None     # [B, rays, Dc] used by every forward while set (depth_jitter takes precedence)
        # decoder MLP arithmetic (ia_render_params.mlp_fmt): 3-term split by default (the op-level API reproduces the fp32 MLP);
        # TriPlaneGenerator sets rt.FMT_F16X1 from its measured error budget, IA_CONV_PRECISION=bf16x3 overrides it (strict mode)
        self.mlp_fmt = rt.FMT_BF16X3

    def _mlp_fmt(self):
        import os
        if os.environ.get('IA_CONV_PRECISION', 'auto') == 'bf16x3':
            return rt.FMT_BF16X3
        return int(getattr(self, 'mlp_fmt', rt.FMT_BF16X3))

    def run_model(self, planes, decoder, sample_coordinates, sample_directions, options):
        """renderer.py:353-363: features of arbitrary 3-D points -> decoder outputs {'rgb', 'sigma'} (shape extraction)."""
        if options.get('density_noise', 0) > 0:
            raise NotImplementedError('density_noise > 0 is a training-time regulariser, not part of the inference path')
        feats = sample_from_planes(self.plane_axes, planes, sample_coordinates, padding_mode='zeros', box_warp=options['box_warp'])
        return decoder(feats, sample_directions)

    def _planes_fp16(self):
        """Whether the caller should hand render_nhwc fp16 planes (TriPlaneGenerator's storage policy; never in strict mode)."""
        import os
        if os.environ.get('IA_CONV_PRECISION', 'auto') == 'bf16x3':
            return False
        return int(getattr(self, 'planes_fmt', rt.FMT_BF16X3)) == rt.FMT_F16X1

    def _draws(self, B, rays, Dc, Df, evaluation, device):
        jit, u = self.depth_jitter, self.importance_u
        # a list pins several consecutive forwards (inversionNet.forward renders twice): one entry is consumed per call
        if isinstance(jit, (list, tuple)):
            jit, self.depth_jitter = (jit[0] if len(jit) else None), (list(jit[1:]) or None)
        else:
            self.depth_jitter = None
        if isinstance(u, (list, tuple)):
            u, self.importance_u = (u[0] if len(u) else None), (list(u[1:]) or None)
        else:
            self.importance_u = None
        if jit is None:
            jit = getattr(self, 'fixed_jitter', None)     # persistent (not consumed) jitter tensor: reproducible video / tests
        if jit is None:
            jit = torch.rand((B, rays, Dc), device=device)
        if evaluation or Df == 0:
            u = None
        elif u is None:
            u = torch.rand((B * rays, Df), device=device)
        return jit.to(device), (u.to(device) if u is not None else None)

    def render_nhwc(self, planes_nhwc, decoder, cam, res, options, evaluation=False):
        """Engine entry: planes [B,PH,PW,96] NHWC, cam [B,25]; returns feat [B,res,res,32], depth, wsum [B,res,res]."""
        if options.get('disparity_space_sampling', False):
            raise NotImplementedError('disparity_space_sampling is not supported (the generator config sets it False)')
        if options.get('density_noise', 0) > 0:
            raise NotImplementedError('density_noise > 0 is a training-time regulariser, not part of the inference path')
        assert options.get('clamp_mode', 'softplus') == 'softplus', 'MipRayMarcher only supports clamp_mode=softplus'
        B = planes_nhwc.shape[0]
        Dc, Df = int(options['depth_resolution']), int(options['depth_resolution_importance'])
        jit, u = self._draws(B, res * res, Dc, Df, evaluation, planes_nhwc.device)
        w1, b1, w2, b2 = _decoder_weights(decoder)
        return rt.render(planes_nhwc, cam, res, Dc, Df, jit, u, float(options['box_warp']), bool(options.get('white_back', False)),
                         w1, b1, w2, b2, mlp_fmt=self._mlp_fmt())

    def forward(self, planes, decoder, ray_origins, ray_directions, rendering_options, evaluation=False):
        """Reference signature: planes [B,3,32,H,W], rays [B,M,3] -> (rgb [B,M,32], depth [B,M,1], weights.sum [B,M,1])."""
        B, n_planes, Cc, H, W = planes.shape
        assert n_planes == 3 and Cc == 32
        M = ray_origins.shape[1]
        res = int(round(math.sqrt(M)))
        assert res * res == M, 'the fused renderer expects a square ray grid'
        planes_nhwc = rt.to_nhwc(planes.reshape(B, n_planes * Cc, H, W))
        Dc, Df = int(rendering_options['depth_resolution']), int(rendering_options['depth_resolution_importance'])
        jit, u = self._draws(B, M, Dc, Df, evaluation, planes.device)
        w1, b1, w2, b2 = _decoder_weights(decoder)
        feat, depth, wsum = rt.render(planes_nhwc, None, res, Dc, Df, jit, u, float(rendering_options['box_warp']),
                                      bool(rendering_options.get('white_back', False)), w1, b1, w2, b2,
                                      rays=(ray_origins, ray_directions), mlp_fmt=self._mlp_fmt())
        return feat.reshape(B, M, 32), depth.reshape(B, M, 1), wsum.reshape(B, M, 1)


ImportanceRenderer = ImportanceRenderer_bsMotion
