"""Oracle restatement of the volume renderer (ray sampler, tri-plane sampling, OSG decoder,
hierarchical importance sampling, MipNeRF-style ray marcher).  TEST INFRASTRUCTURE ONLY.

The two random draws of the reference (coarse depth jitter ``torch.rand_like``,
renderer.py:406, and importance ``torch.rand``, renderer.py:453) are explicit inputs here.
"""
import torch
import torch.nn.functional as F

from . import stylegan2 as sg


def ray_sampler_zxc(cam2world, intrinsics, resolution):
    """volumetric_rendering/ray_sampler.py:70-107: K scaled by res, pixel grid (i,j,1) with no
    half-pixel offset, K^-1, rotate by c2w, normalise; origin = c2w[:3,3]."""
    N = cam2world.shape[0]
    K = intrinsics.clone()
    K[:, :2] *= resolution
    lin = torch.linspace(0, resolution - 1, resolution)
    yy, xx = torch.meshgrid(lin, lin, indexing='ij')
    homo = torch.stack((xx, yy, torch.ones_like(xx)), -1)  # [H, W, 3] = (col, row, 1)
    origins, dirs = [], []
    for n in range(N):
        K_inv = torch.linalg.inv(K[n])
        d = (K_inv[None, ...] @ homo[..., None])[:, :, :, 0]
        d = (cam2world[n][None, :3, :3] @ d[..., None])[:, :, :, 0]
        d = F.normalize(d, dim=-1)
        o = cam2world[n][:3, -1].expand(d.shape)
        dirs.append(d.reshape(-1, 3))
        origins.append(o.reshape(-1, 3))
    return torch.stack(origins, 0), torch.stack(dirs, 0)


def plane_axes_inv():
    """volumetric_rendering/renderer.py:30-48."""
    planes = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                           [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                           [[0, 0, 1], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)
    return torch.linalg.inv(planes)


def sample_from_planes(planes, coordinates, box_warp):
    """renderer.py:51-65,85-97: project onto (x,y),(x,z),(z,x); bilinear, zeros padding,
    align_corners=False."""
    N, n_planes, C, H, W = planes.shape
    M = coordinates.shape[1]
    feats = planes.reshape(N * n_planes, C, H, W)
    coords = (2 / box_warp) * coordinates
    inv = plane_axes_inv()
    cc = coords.unsqueeze(1).expand(-1, n_planes, -1, -1).reshape(N * n_planes, M, 3)
    iv = inv.unsqueeze(0).expand(N, -1, -1, -1).reshape(N * n_planes, 3, 3)
    proj = torch.bmm(cc, iv)[..., :2].unsqueeze(1)
    out = F.grid_sample(feats, proj.float(), mode='bilinear', padding_mode='zeros', align_corners=False)
    return out.permute(0, 3, 2, 1).reshape(N, n_planes, M, C)


def osg_decoder(sd, sampled_features):
    """training_avatar_texture/triplane_v20.py:426-438."""
    x = sampled_features.mean(1)
    N, M, C = x.shape
    x = x.reshape(N * M, C)
    x = sg.fully_connected(x, sd['net.0.weight'], sd['net.0.bias'])
    x = F.softplus(x)
    x = sg.fully_connected(x, sd['net.2.weight'], sd['net.2.bias'])
    x = x.reshape(N, M, -1)
    rgb = torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001
    return rgb, x[..., 0:1]


def ray_march(colors, densities, depths, white_back=False):
    """volumetric_rendering/ray_marcher.py:25-57 (clamp_mode softplus)."""
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    colors_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    dens_mid = (densities[:, :, :-1] + densities[:, :, 1:]) / 2
    depths_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    dens_mid = F.softplus(dens_mid - 1)
    alpha = 1 - torch.exp(-dens_mid * deltas)
    alpha_shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2)
    weights = alpha * torch.cumprod(alpha_shifted, -2)[:, :, :-1]
    rgb = torch.sum(weights * colors_mid, -2)
    wtot = weights.sum(2)
    depth = torch.sum(weights * depths_mid, -2) / wtot
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))
    if white_back:
        rgb = rgb + 1 - wtot
    rgb = rgb * 2 - 1
    return rgb, depth, weights


def sample_stratified(ray_origins, ray_start, ray_end, depth_resolution, jitter):
    """renderer.py:384-408, linear (non-disparity) branch; ``jitter`` replaces rand_like."""
    N, M, _ = ray_origins.shape
    d = torch.linspace(ray_start, ray_end, depth_resolution).reshape(1, 1, depth_resolution, 1).repeat(N, M, 1, 1)
    delta = (ray_end - ray_start) / (depth_resolution - 1)
    return d + jitter * delta


def sample_pdf(bins, weights, n_importance, det, u=None, eps=1e-5):
    """renderer.py:427-469; ``u`` replaces torch.rand when det is False."""
    n_rays, n_samples = weights.shape
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    if det:
        u = torch.linspace(0, 1, n_importance).expand(n_rays, n_importance)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_samples)
    idx = torch.stack([below, above], -1).view(n_rays, 2 * n_importance)
    cdf_g = torch.gather(cdf, 1, idx).view(n_rays, n_importance, 2)
    bins_g = torch.gather(bins, 1, idx).view(n_rays, n_importance, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom[denom < eps] = 1
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])


def sample_importance(z_vals, weights, n_importance, det, u=None):
    """renderer.py:410-425: max-pool/avg-pool smoothing, +0.01, inverse-CDF sampling."""
    B, R, S, _ = z_vals.shape
    z = z_vals.reshape(B * R, S)
    w = weights.reshape(B * R, -1)
    w = F.max_pool1d(w.unsqueeze(1).float(), 2, 1, padding=1)
    w = F.avg_pool1d(w, 2, 1).squeeze(1)
    w = w + 0.01
    z_mid = 0.5 * (z[:, :-1] + z[:, 1:])
    out = sample_pdf(z_mid, w[:, 1:-1], n_importance, det, u)
    return out.reshape(B, R, n_importance, 1)


def unify_samples(d1, c1, s1, d2, c2, s2):
    """renderer.py:372-382."""
    d = torch.cat([d1, d2], dim=-2)
    c = torch.cat([c1, c2], dim=-2)
    s = torch.cat([s1, s2], dim=-2)
    _, idx = torch.sort(d, dim=-2)
    d = torch.gather(d, -2, idx)
    c = torch.gather(c, -2, idx.expand(-1, -1, -1, c.shape[-1]))
    s = torch.gather(s, -2, idx.expand(-1, -1, -1, 1))
    return d, c, s


def importance_renderer(decoder_sd, planes, ray_origins, ray_directions, options, jitter,
                        evaluation=True, u=None):
    """ImportanceRenderer_bsMotion.forward, renderer.py:309-351."""
    dist = torch.norm(ray_origins, dim=-1).mean().item()
    ray_start, ray_end = dist - 0.45, dist + 0.6
    Dc = options['depth_resolution']
    depths_c = sample_stratified(ray_origins, ray_start, ray_end, Dc, jitter)
    B, R, S, _ = depths_c.shape

    def run_model(depths, n):
        pts = (ray_origins.unsqueeze(-2) + depths * ray_directions.unsqueeze(-2)).reshape(B, -1, 3)
        feats = sample_from_planes(planes, pts, options['box_warp'])
        rgb, sigma = osg_decoder(decoder_sd, feats)
        return rgb.reshape(B, R, n, rgb.shape[-1]), sigma.reshape(B, R, n, 1)

    colors_c, dens_c = run_model(depths_c, S)
    Df = options['depth_resolution_importance']
    white_back = options.get('white_back', False)
    if Df > 0:
        _, _, weights = ray_march(colors_c, dens_c, depths_c, white_back)
        depths_f = sample_importance(depths_c, weights, Df, det=evaluation, u=u)
        colors_f, dens_f = run_model(depths_f, Df)
        d, c, s = unify_samples(depths_c, colors_c, dens_c, depths_f, colors_f, dens_f)
        rgb, depth, weights = ray_march(c, s, d, white_back)
    else:
        rgb, depth, weights = ray_march(colors_c, dens_c, depths_c, white_back)
    return rgb, depth, weights.sum(2)
