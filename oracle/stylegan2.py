"""Oracle restatement of the StyleGAN2 generator blocks used by the Next3D++ generator.
Functional style over a flat state dict (reference key names).  TEST INFRASTRUCTURE ONLY.

Follows reference training_avatar_texture/networks_stylegan2_new.py (cond_list / return_list /
feat_conditions variant) and training/networks_stylegan2.py (stock blocks used by the
super-resolution module); both share the same layer arithmetic.
"""
import math

import numpy as np
import torch

from . import ops


def sub(sd, prefix):
    """View of a state dict below ``prefix`` (with trailing dot stripped)."""
    p = prefix + '.'
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def fully_connected(x, weight, bias=None, activation='linear', lr_multiplier=1.0):
    """networks_stylegan2_new.py:96-127: runtime gains w*lr/sqrt(in), b*lr."""
    w = weight * (lr_multiplier / math.sqrt(weight.shape[1]))
    b = bias
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    if activation == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    x = x.matmul(w.t())
    return ops.bias_act(x, b, act=activation)


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    """networks_stylegan2_new.py:28-29."""
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def mapping_network(sd, z, c, num_ws, num_layers=2, truncation_psi=1.0, truncation_cutoff=None,
                    lr_multiplier=0.01):
    """networks_stylegan2_new.py:233-268 (no EMA update)."""
    x = normalize_2nd_moment(z.float())
    if 'embed.weight' in sd:
        y = normalize_2nd_moment(fully_connected(c.float(), sd['embed.weight'], sd['embed.bias']))
        x = torch.cat([x, y], dim=1)
    for i in range(num_layers):
        x = fully_connected(x, sd[f'fc{i}.weight'], sd[f'fc{i}.bias'], activation='lrelu',
                            lr_multiplier=lr_multiplier)
    x = x.unsqueeze(1).repeat(1, num_ws, 1)
    if truncation_psi != 1:
        w_avg = sd['w_avg']
        if truncation_cutoff is None:
            x = w_avg.lerp(x, truncation_psi)
        else:
            x[:, :truncation_cutoff] = w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
    return x


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    """networks_stylegan2_new.py:34-91 (fp32 branches)."""
    B = x.shape[0]
    O, I, kh, kw = weight.shape
    w = weight.unsqueeze(0) * styles.reshape(B, 1, I, 1, 1)
    dcoefs = None
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if not fused_modconv:
        x = x * styles.reshape(B, I, 1, 1)
        x = ops.conv2d_resample(x, weight, f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = x * dcoefs.reshape(B, O, 1, 1) + noise
        elif demodulate:
            x = x * dcoefs.reshape(B, O, 1, 1)
        elif noise is not None:
            x = x + noise
        return x
    if demodulate:
        w = w * dcoefs.reshape(B, O, 1, 1, 1)
    x = x.reshape(1, B * I, *x.shape[2:])
    w = w.reshape(B * O, I, kh, kw)
    x = ops.conv2d_resample(x, w, f=resample_filter, up=up, padding=padding, groups=B, flip_weight=flip_weight)
    x = x.reshape(B, O, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x


def synthesis_layer(sd, x, w, up=1, noise_mode='const', fused_modconv=True, gain=1.0, conv_clamp=None):
    """networks_stylegan2_new.py:311-330."""
    styles = fully_connected(w, sd['affine.weight'], sd['affine.bias'])
    noise = None
    if noise_mode == 'const' and 'noise_const' in sd:
        noise = sd['noise_const'] * sd['noise_strength']
    elif noise_mode == 'random':
        raise NotImplementedError('oracle supports noise_mode const|none')
    x = modulated_conv2d(x, sd['weight'], styles, noise=noise, up=up, padding=1,
                         resample_filter=sd['resample_filter'], flip_weight=(up == 1),
                         fused_modconv=fused_modconv)
    act_gain = ops.SQRT2 * gain
    act_clamp = conv_clamp * gain if conv_clamp is not None else None
    return ops.bias_act(x, sd['bias'], act='lrelu', gain=act_gain, clamp=act_clamp)


def torgb_layer(sd, x, w, fused_modconv=True, conv_clamp=None):
    """networks_stylegan2_new.py:353-357."""
    I = sd['weight'].shape[1]
    styles = fully_connected(w, sd['affine.weight'], sd['affine.bias']) * (1 / math.sqrt(I))
    x = modulated_conv2d(x, sd['weight'], styles, demodulate=False, fused_modconv=fused_modconv)
    return ops.bias_act(x, sd['bias'], clamp=conv_clamp)


def synthesis_block(sd, x, img, ws, condition=None, noise_mode='const', fused_modconv=True,
                    conv_clamp=None, is_last=False):
    """networks_stylegan2_new.py:417-467, skip architecture, fp32."""
    widx = 0
    if 'const' in sd:
        x = sd['const'].unsqueeze(0).repeat(ws.shape[0], 1, 1, 1)
        x = synthesis_layer(sub(sd, 'conv1'), x, ws[:, widx], noise_mode=noise_mode,
                            fused_modconv=fused_modconv, conv_clamp=conv_clamp)
        widx += 1
    else:
        x = synthesis_layer(sub(sd, 'conv0'), x, ws[:, widx], up=2, noise_mode=noise_mode,
                            fused_modconv=fused_modconv, conv_clamp=conv_clamp)
        widx += 1
        if condition is not None:  # CS-SFT, :448-452
            half = x.shape[1] // 2
            x = torch.cat([x[:, :half], x[:, half:] * condition[0] + condition[1]], dim=1)
        x = synthesis_layer(sub(sd, 'conv1'), x, ws[:, widx], noise_mode=noise_mode,
                            fused_modconv=fused_modconv, conv_clamp=conv_clamp)
        widx += 1
    if img is not None:
        img = ops.upsample2d(img, sd['resample_filter'])
    y = torgb_layer(sub(sd, 'torgb'), x, ws[:, widx], fused_modconv=fused_modconv, conv_clamp=conv_clamp)
    img = img + y if img is not None else y
    return x, img


def block_resolutions(sd):
    res = sorted({int(k.split('.')[0][1:]) for k in sd.keys() if k.startswith('b')})
    return res


def synthesis_network(sd, ws, cond_list=None, return_list=False, feat_conditions=None,
                      out_res=(32, 256), noise_mode='const', fused_modconv=True):
    """networks_stylegan2_new.py:509-548."""
    resolutions = block_resolutions(sd)
    img_res_log2 = int(np.log2(resolutions[-1]))
    x = img = None
    x_list = []
    start_layer = int(np.log2(out_res[0])) - 2
    end_layer = (img_res_log2 - 2) if len(out_res) == 1 else (int(np.log2(out_res[1])) - 2)
    w_idx = 0
    for index, res in enumerate(resolutions):
        bsd = sub(sd, f'b{res}')
        n_conv = 1 if 'const' in bsd else 2
        cur_ws = ws[:, w_idx:w_idx + n_conv + 1]
        w_idx += n_conv
        cond_feat = feat_conditions[res] if (feat_conditions is not None and res in feat_conditions) else None
        x, img = synthesis_block(bsd, x, img, cur_ws, cond_feat, noise_mode=noise_mode,
                                 fused_modconv=fused_modconv)
        if index >= start_layer:
            if return_list:
                if index == start_layer:
                    x_list.append(img.clone())
                x_list.append(x.clone())
            if cond_list is not None:
                if index == start_layer:
                    a = cond_list[0][:, -1:]
                    img = cond_list[0][:, :-1] * a + img * (1 - a)
                if index < end_layer:
                    cnd = cond_list[1 + index - start_layer]
                    a = cnd[:, -1:]
                    x = cnd[:, :-1] * a + x * (1 - a)
    if return_list:
        x_list.append(img)
        return x_list
    return img


def superresolution_8xdc(sd, rgb, x, ws, noise_mode='none', sr_antialias=True, input_resolution=128):
    """training_avatar_texture/superresolution.py:263-289 with the stock SynthesisBlock
    (training/networks_stylegan2.py:417ff); clamp 256 because sr_num_fp16_res>0, fp32 on CPU
    (SURVEY appendix B)."""
    ws = ws[:, -1:, :].repeat(1, 3, 1)
    if x.shape[-1] != input_resolution:
        size = (input_resolution, input_resolution)
        x = torch.nn.functional.interpolate(x, size=size, mode='bilinear', align_corners=False, antialias=sr_antialias)
        rgb = torch.nn.functional.interpolate(rgb, size=size, mode='bilinear', align_corners=False, antialias=sr_antialias)
    x, rgb = synthesis_block(sub(sd, 'block0'), x, rgb, ws, noise_mode=noise_mode, conv_clamp=256)
    x, rgb = synthesis_block(sub(sd, 'block1'), x, rgb, ws, noise_mode=noise_mode, conv_clamp=256)
    return rgb
