"""Oracle restatement of the Next3D++ TriPlaneGenerator forward
(reference training_avatar_texture/triplane_v20.py).  TEST INFRASTRUCTURE ONLY."""
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F

from . import renderer as rd
from . import stylegan2 as sg

BBOX_256 = (57, 185, 64, 192)  # triplane_v20.py:114


def flood_fill_fixed_range(img, lo=0.0, up=254.0):
    """Numpy restatement of ``cv2.floodFill(img, mask, (0,0), 255, lo, up, FLOODFILL_FIXED_RANGE)``
    on a float32 single-channel image (OpenCV 4.x floodfill.cpp, Diff32fC1, 4-connectivity):
    a pixel is filled when it is 4-connected to the seed through pixels p with
    ``-lo <= p - seed_value <= up`` (original values).  Returns the boolean filled mask."""
    h, w = img.shape
    seed = np.float32(img[0, 0])
    d = img.astype(np.float32) - seed
    passable = (d >= np.float32(-lo)) & (d <= np.float32(up))
    filled = np.zeros((h, w), dtype=bool)
    filled[0, 0] = True  # the seed itself is always filled
    q = deque([(0, 0)])
    while q:
        y, x = q.popleft()
        for ny, nx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
            if 0 <= ny < h and 0 <= nx < w and passable[ny, nx] and not filled[ny, nx]:
                filled[ny, nx] = True
                q.append((ny, nx))
    return filled


def fill_mouth(images):
    """volumetric_rendering/renderer.py:716-741 with blur_mouth_edge=False: mouth mask =
    (255 - floodfilled(alpha*255)) / 255, i.e. 0 where the background flood reaches and
    1-alpha elsewhere; result = clip(alpha + mask, 0, 1)."""
    masks = []
    for image in images:
        im = image[0].cpu().numpy().astype(np.float32) * np.float32(255.0)
        filled = flood_fill_fixed_range(im)
        out = im.copy()
        out[filled] = np.float32(255.0)
        masks.append(torch.tensor(np.float32(255.0) - out).to(torch.float32).unsqueeze(0) / 255.0)
    masks = torch.stack(masks, 0)
    return (images + masks).clip(0, 1), masks


def rasterize(texture_feats, uvcoords_image, static_feats, bbox_256=BBOX_256):
    """triplane_v20.py:317-339."""
    uv = uvcoords_image.float()
    grid, alpha = uv[..., :2], uv[..., 2:].permute(0, 3, 1, 2)
    full_alpha, mouth = fill_mouth(alpha.clone())
    upper = mouth.clone()
    upper[:, :, :87] = 0
    upper_alpha = torch.clamp(alpha + upper, min=0, max=1)
    outs = []
    for idx, tex in enumerate(texture_feats):
        res = tex.shape[2]
        bbox = [round(i * res / 256) for i in bbox_256]
        ri = F.grid_sample(tex, grid, align_corners=False)
        rf = F.interpolate(ri, size=(res, res), mode='bilinear', antialias=True)
        a = F.interpolate(alpha, size=(res, res), mode='bilinear', antialias=True)
        st = F.interpolate(static_feats[idx][:, :, bbox[0]:bbox[1], bbox[2]:bbox[3]], size=(res, res),
                           mode='bilinear', antialias=True)
        ua = F.interpolate(upper_alpha, size=(res, res), mode='bilinear', antialias=True)
        outs.append(torch.cat([rf * a + st * (1 - a), ua], dim=1))
    return outs, full_alpha, mouth


def mapping(sd, z, c, rendering_kwargs, truncation_psi=1.0, truncation_cutoff=None, c_dim=25, num_ws=14):
    """triplane_v20.py:64-69."""
    if rendering_kwargs['c_gen_conditioning_zero']:
        c = torch.zeros_like(c)
    c = c[:, :c_dim]
    return sg.mapping_network(sg.sub(sd, 'backbone.mapping'), z, c * rendering_kwargs.get('c_scale', 0),
                              num_ws=num_ws, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)


def _split_static(static_feats):
    """triplane_v20.py:109-112: keep plane 0 of the 96-channel images for the rasterizer."""
    sf = list(static_feats)
    B = sf[-1].shape[0]
    plane = sf[-1].view(B, 3, 32, sf[-1].shape[-2], sf[-1].shape[-1])
    sf[0] = sf[0].view(B, 3, 32, sf[0].shape[-2], sf[0].shape[-1])[:, 0]
    sf[-1] = plane[:, 0]
    return sf, plane


def _stitch_render_sr(sd, ws, c, uv, texture_feats, static_feats_raw, rendering_kwargs, jitter, evaluation, u,
                      neural_rendering_resolution, stages):
    cam = c[:, -25:]
    c2w = cam[:, :16].view(-1, 4, 4)
    K = cam[:, 16:25].view(-1, 3, 3)
    ray_o, ray_d = rd.ray_sampler_zxc(c2w, K, neural_rendering_resolution)
    static_feats, static_plane = _split_static(static_feats_raw)
    rendering_images, full_alpha, _ = rasterize(texture_feats, uv, static_feats)
    stitch = sg.synthesis_network(sg.sub(sd, 'face_backbone.synthesis'), ws, cond_list=rendering_images,
                                  return_list=False)
    b0, b1, b2, b3 = BBOX_256
    stitch_ = torch.zeros_like(stitch)
    alpha_ = torch.zeros_like(full_alpha)
    stitch_[:, :, b0:b1, b2:b3] = F.interpolate(stitch, size=(128, 128), mode='bilinear', antialias=True)
    alpha_[:, :, b0:b1, b2:b3] = F.interpolate(full_alpha, size=(128, 128), mode='bilinear', antialias=True)
    alpha3 = torch.cat((alpha_, torch.zeros_like(alpha_), torch.zeros_like(alpha_)), 1).unsqueeze(2)
    stitch3 = torch.cat((stitch_, torch.zeros_like(stitch_), torch.zeros_like(stitch_)), 1).view(*static_plane.shape)
    planes = stitch3 * alpha3 + static_plane * (1 - alpha3)
    feat, depth, wsum = rd.importance_renderer(sg.sub(sd, 'decoder'), planes, ray_o, ray_d, rendering_kwargs,
                                               jitter, evaluation=evaluation, u=u)
    H = W = neural_rendering_resolution
    N = ws.shape[0]
    feature_image = feat.permute(0, 2, 1).reshape(N, feat.shape[-1], H, W).contiguous()
    depth_image = depth.permute(0, 2, 1).reshape(N, 1, H, W)
    rgb_image = feature_image[:, :3]
    sr = sg.superresolution_8xdc(sg.sub(sd, 'superresolution'), rgb_image, feature_image, ws,
                                 noise_mode=rendering_kwargs['superresolution_noise_mode'],
                                 sr_antialias=rendering_kwargs['sr_antialias'])
    out = {'image': sr, 'image_raw': rgb_image, 'image_depth': depth_image}
    if stages:
        out.update({'feature_image': feature_image, 'triplane': planes, 'rendering_images': rendering_images,
                    'full_alpha': full_alpha, 'rendering_stitch': stitch, 'weights_sum': wsum,
                    'ray_origins': ray_o, 'ray_directions': ray_d})
    return out


def synthesis(sd, ws, c, uvcoords_image, rendering_kwargs, jitter, evaluation=True, u=None,
              neural_rendering_resolution=128, stages=False):
    """triplane_v20.py:89-150 with noise_mode='const'."""
    texture_feats = sg.synthesis_network(sg.sub(sd, 'texture_backbone.synthesis'), ws, return_list=True)
    static_raw = sg.synthesis_network(sg.sub(sd, 'backbone.synthesis'), ws, return_list=True)
    out = _stitch_render_sr(sd, ws, c, uvcoords_image, texture_feats, static_raw, rendering_kwargs, jitter,
                            evaluation, u, neural_rendering_resolution, stages)
    if stages:
        out['texture_feats'] = texture_feats
        out['static_feats'] = static_raw
    return out


def synthesis_with_texture(sd, ws, texture_feats, c, uvcoords_image, rendering_kwargs, jitter,
                           static_feats=None, evaluation=True, u=None, neural_rendering_resolution=128,
                           stages=False):
    """triplane_v20.py:152-244."""
    if static_feats is None:
        static_feats = sg.synthesis_network(sg.sub(sd, 'backbone.synthesis'), ws, return_list=True)
    return _stitch_render_sr(sd, ws, c, uvcoords_image, texture_feats, static_feats, rendering_kwargs, jitter,
                             evaluation, u, neural_rendering_resolution, stages)


def sample_mixed(sd, coordinates, ws, uvcoords_image, rendering_kwargs):
    """triplane_v20.py:373-402 (= ``sample`` after the mapping, :341-371): decoder outputs at arbitrary 3-D points, from the
    blended tri-planes of the frame; ``run_model`` (renderer.py:353-363) without density noise.  Returns (rgb, sigma)."""
    texture_feats = sg.synthesis_network(sg.sub(sd, 'texture_backbone.synthesis'), ws, return_list=True)
    static_raw = sg.synthesis_network(sg.sub(sd, 'backbone.synthesis'), ws, return_list=True)
    static_feats, static_plane = _split_static(static_raw)
    rendering_images, full_alpha, _ = rasterize(texture_feats, uvcoords_image, static_feats)
    stitch = sg.synthesis_network(sg.sub(sd, 'face_backbone.synthesis'), ws, cond_list=rendering_images, return_list=False)
    b0, b1, b2, b3 = BBOX_256
    stitch_, alpha_ = torch.zeros_like(stitch), torch.zeros_like(full_alpha)
    stitch_[:, :, b0:b1, b2:b3] = F.interpolate(stitch, size=(128, 128), mode='bilinear', antialias=True)
    alpha_[:, :, b0:b1, b2:b3] = F.interpolate(full_alpha, size=(128, 128), mode='bilinear', antialias=True)
    alpha3 = torch.cat((alpha_, torch.zeros_like(alpha_), torch.zeros_like(alpha_)), 1).unsqueeze(2)
    stitch3 = torch.cat((stitch_, torch.zeros_like(stitch_), torch.zeros_like(stitch_)), 1).view(*static_plane.shape)
    planes = stitch3 * alpha3 + static_plane * (1 - alpha3)
    feats = rd.sample_from_planes(planes, coordinates, rendering_kwargs['box_warp'])
    return rd.osg_decoder(sg.sub(sd, 'decoder'), feats)


def visualize_mesh_condition(uvcoords_image):
    """triplane_v20.py:71-87 with to_imgs=True, up to the PIL conversion: uint8 [B,3,H,W]."""
    uv = uvcoords_image.clone().permute(0, 3, 1, 2)
    full_alpha, _ = fill_mouth(uv[:, 2:].clone())
    uv[full_alpha.expand(-1, 3, -1, -1) == 0] = -1
    return ((uv + 1) * 127.5).to(dtype=torch.uint8)
