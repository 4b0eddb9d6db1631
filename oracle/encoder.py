"""CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py) of the inversion encoder that feeds the generator:
e4e (IR-SE50 + FPN + GradualStyleBlocks), the texture UNet and the tri-plane SFT UNet with their ConvGRU decoders,
and ``inversionNet.{encode, get_unet_uvinput, AR_eval_forward}``.

Functional restatement over a state-dict (same key names as the reference modules); every function cites the reference
file:line it follows (paths relative to /root/reference/encoder_inversion/models).  Convolutions / interpolation /
grid_sample are PyTorch ATen in both the reference and here (third-party arithmetic, oracle/__init__.py).

BatchNorm mode: eval_seq.py:92-97 puts the whole inversionNet in train mode and re-``eval()``s only the two UNets'
``input_layer`` / ``body``; every BatchNorm therefore carries an explicit ``training`` flag here (batch statistics,
biased variance, eps 1e-5; running statistics are not updated by the oracle because they do not feed the output).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import stylegan2 as o_sg

BN_EPS = 1e-5


def sub(sd, prefix):
    return o_sg.sub(sd, prefix)


# ---- helpers.py --------------------------------------------------------------------------------------------
def get_blocks50():
    """helpers.py:30-41: (in_channel, depth, stride) of the 24 IR-SE50 units."""
    def block(i, d, n):
        return [(i, d, 2)] + [(d, d, 1)] * (n - 1)
    return block(64, 64, 3) + block(64, 128, 4) + block(128, 256, 14) + block(256, 512, 3)


def batch_norm(sd, x, training):
    """torch.nn.BatchNorm2d forward: batch statistics (biased variance) in train mode, running statistics in eval."""
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
    else:
        mean, var = sd['running_mean'], sd['running_var']
    scale = sd['weight'] / torch.sqrt(var + BN_EPS)
    shift = sd['bias'] - mean * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def prelu(x, weight):
    return torch.where(x >= 0, x, x * weight.view(1, -1, 1, 1))


def se_module(sd, x):
    """helpers.py:62-80: squeeze-excite with two bias-free 1x1 convolutions."""
    s = x.mean(dim=(2, 3), keepdim=True)
    s = F.relu(F.conv2d(s, sd['fc1.weight']))
    s = torch.sigmoid(F.conv2d(s, sd['fc2.weight']))
    return x * s


def bottleneck_ir_se(sd, x, in_c, depth, stride, training):
    """helpers.py:102-124."""
    if in_c == depth:
        shortcut = x[:, :, ::stride, ::stride]                       # MaxPool2d(1, stride)
    else:
        shortcut = F.conv2d(x, sd['shortcut_layer.0.weight'], stride=stride)
        shortcut = batch_norm(sub(sd, 'shortcut_layer.1'), shortcut, training)
    r = batch_norm(sub(sd, 'res_layer.0'), x, training)
    r = F.conv2d(r, sd['res_layer.1.weight'], padding=1)
    r = prelu(r, sd['res_layer.2.weight'])
    r = F.conv2d(r, sd['res_layer.3.weight'], stride=stride, padding=1)
    r = batch_norm(sub(sd, 'res_layer.4'), r, training)
    r = se_module(sub(sd, 'res_layer.5'), r)
    return r + shortcut


def ir_se50_trunk(sd, x, taps, training):
    """input_layer + body with feature taps after the given body indices (e4e.py:105-117, unet_encoders.py:201-212)."""
    x = F.conv2d(x, sd['input_layer.0.weight'], padding=1)
    x = batch_norm(sub(sd, 'input_layer.1'), x, training)
    x = prelu(x, sd['input_layer.2.weight'])
    feats = {}
    for i, (in_c, depth, stride) in enumerate(get_blocks50()):
        x = bottleneck_ir_se(sub(sd, f'body.{i}'), x, in_c, depth, stride, training)
        if i in taps:
            feats[i] = x
    return x, feats


# ---- e4e.py ------------------------------------------------------------------------------------------------
def gradual_style_block(sd, x, spatial):
    """e4e.py:22-46: log2(spatial) stride-2 convolutions with LeakyReLU(0.01), then an equalised linear layer."""
    n = int(np.log2(spatial))
    for j in range(n):
        x = F.leaky_relu(F.conv2d(x, sd[f'convs.{2 * j}.weight'], sd[f'convs.{2 * j}.bias'], stride=2, padding=1), 0.01)
    x = x.reshape(-1, x.shape[1])
    return o_sg.fully_connected(x, sd['linear.weight'], sd['linear.bias'])


def upsample_add(x, y):
    """e4e.py:48-65."""
    return F.interpolate(x, size=y.shape[-2:], mode='bilinear', align_corners=True) + y


def encoder4editing(sd, x, n_styles=14, training=False):
    """e4e.py:105-134."""
    c3, f = ir_se50_trunk(sd, x, taps=(6, 20, 23), training=training)
    c1, c2 = f[6], f[20]
    coarse_ind, middle_ind = 3, 7

    def style(i, feat):
        spatial = 16 if i < coarse_ind else (32 if i < middle_ind else 64)
        return gradual_style_block(sub(sd, f'styles.{i}'), feat, spatial)
    w0 = style(0, c3)
    w = w0.unsqueeze(1).repeat(1, n_styles, 1)
    features = c3
    p2 = None
    for i in range(1, n_styles):
        if i == coarse_ind:
            p2 = upsample_add(c3, F.conv2d(c2, sd['latlayer1.weight'], sd['latlayer1.bias']))
            features = p2
        elif i == middle_ind:
            features = upsample_add(p2, F.conv2d(c1, sd['latlayer2.weight'], sd['latlayer2.bias']))
        w[:, i] = w[:, i] + style(i, features)
    return w


def face_pool(x, res=256):
    """AdaptiveAvgPool2d((256,256)) applied only when the input is not already res wide (uvnet.py:108-109)."""
    return F.adaptive_avg_pool2d(x, (res, res)) if x.shape[-1] != res else x


def encode(sd, x, n_styles=14, training=False):
    """inversionNet.encode, uvnet.py:107-115.  sd: inversionNet state-dict."""
    codes = encoder4editing(sub(sd, 'encoder'), face_pool(x), n_styles, training)
    return codes + sd['latent_avg'].reshape(1, 1, -1)


# ---- unet_encoders.py --------------------------------------------------------------------------------------
def conv_gru_step(sd, x, h):
    """unet_encoders.py:27-32."""
    Cc = x.shape[1]
    rz = torch.sigmoid(F.conv2d(torch.cat([x, h], dim=1), sd['ih.0.weight'], sd['ih.0.bias'], padding=1))
    r, z = rz.split(Cc, dim=1)
    c = torch.tanh(F.conv2d(torch.cat([x, r * h], dim=1), sd['hh.0.weight'], sd['hh.0.bias'], padding=1))
    return (1 - z) * h + z * c


def conv_gru(sd, x, h):
    """unet_encoders.py:34-49 with seq2seq=False: x [B,T,C,H,W] -> last hidden state (returned as output and state)."""
    if h is None:
        h = torch.zeros_like(x[:, 0])
    for t in range(x.shape[1]):
        h = conv_gru_step(sd, x[:, t], h)
    return h, h


def double_conv(sd, x, training):
    """unet_encoders.py:52-67: BN -> conv3x3 -> PReLU -> conv3x3 -> PReLU -> PReLU."""
    x = batch_norm(sub(sd, 'double_conv.0'), x, training)
    x = prelu(F.conv2d(x, sd['double_conv.1.weight'], sd['double_conv.1.bias'], padding=1), sd['double_conv.2.weight'])
    x = prelu(F.conv2d(x, sd['double_conv.3.weight'], sd['double_conv.3.bias'], padding=1), sd['double_conv.4.weight'])
    return prelu(x, sd['double_conv.5.weight'])


def recurrent_up(sd, x1, x2, T, r, upscale, training):
    """unet_encoders.py:85-98."""
    if upscale > 1:
        x1 = F.pixel_shuffle(x1, upscale)
    x = double_conv(sub(sd, 'conv'), torch.cat([x2, x1], dim=1), training)
    return conv_gru(sub(sd, 'conv_gru'), x.unflatten(0, (-1, T)), r)


def _unet_trunk_and_decoder(sd, x, r_list, bn_trunk_training, bn_decoder_training, res=256):
    """Shared part of the two UNets (unet_encoders.py:193-232 / 304-345): returns the four decoder states."""
    if x.dim() == 5:
        T = x.shape[1]
        x = x.flatten(0, 1)
    else:
        T = 1
    x = face_pool(x, res)
    x, f = ir_se50_trunk(sd, x, taps=(2, 6, 20, 21), training=bn_trunk_training)
    c0, c1, c2, c3 = f[2], f[6], f[20], f[21]
    r_list = [None] * 4 if r_list is None else list(r_list)

    def rep(t):
        return t.unsqueeze(1).expand(-1, T, -1, -1, -1).flatten(0, 1)
    t1, r_list[0] = recurrent_up(sub(sd, 'up1'), x, c3, T, r_list[0], 1, bn_decoder_training)
    t2, r_list[1] = recurrent_up(sub(sd, 'up2'), rep(t1), c2, T, r_list[1], 2, bn_decoder_training)
    t3, r_list[2] = recurrent_up(sub(sd, 'up3'), rep(t2), c1, T, r_list[2], 2, bn_decoder_training)
    t4, r_list[3] = recurrent_up(sub(sd, 'up4'), rep(t3), c0, T, r_list[3], 2, bn_decoder_training)
    return (t1, t2, t3, t4), r_list


def texture_unet(sd, x, r_list=None, bn_trunk_training=False, bn_decoder_training=True):
    """TriPlanefeat_Encoder.forward (use_gru=True), unet_encoders.py:193-232 -> ([4 offsets], r_list)."""
    (t1, t2, t3, t4), r_list = _unet_trunk_and_decoder(sd, x, r_list, bn_trunk_training, bn_decoder_training)
    outs = [F.conv2d(t2, sd['outconv0.weight'], sd['outconv0.bias']), F.conv2d(t2, sd['outconv1.weight'], sd['outconv1.bias']),
            F.conv2d(t3, sd['outconv2.weight'], sd['outconv2.bias']), F.conv2d(t4, sd['outconv3.weight'], sd['outconv3.bias'])]
    return outs, r_list


def _sft_head(sd, name, x):
    y = F.leaky_relu(F.conv2d(x, sd[f'{name}.0.weight'], sd[f'{name}.0.bias'], padding=1), 0.2)
    return F.conv2d(y, sd[f'{name}.2.weight'], sd[f'{name}.2.bias'], padding=1)


def triplane_unet(sd, x, r_list=None, bn_trunk_training=False, bn_decoder_training=True):
    """TriPlaneSFTfeat_Encoder.forward (use_gru=True), unet_encoders.py:304-345 -> ({res: stack(scale, shift)}, r_list)."""
    (t1, t2, t3, t4), r_list = _unet_trunk_and_decoder(sd, x, r_list, bn_trunk_training, bn_decoder_training)
    t5 = F.pixel_shuffle(t4, 2)
    t5 = prelu(F.conv2d(t5, sd['final_head.0.weight'], sd['final_head.0.bias'], padding=1), sd['final_head.1.weight'])
    t5 = prelu(F.conv2d(t5, sd['final_head.2.weight'], sd['final_head.2.bias'], padding=1), sd['final_head.3.weight'])
    out = {}
    for res, t in zip((16, 32, 64, 128, 256), (t1, t2, t3, t4, t5)):
        out[res] = torch.stack([_sft_head(sd, f'condition_scale{res}', t), _sft_head(sd, f'condition_shift{res}', t)])
    return out, r_list


# ---- uvnet.py ----------------------------------------------------------------------------------------------
def get_unet_uvinput(sd, uv, delta_x):
    """uvnet.py:117-121."""
    uv_gttex, uv_pverts = uv.split(3, dim=1)
    uv_delta = F.grid_sample(delta_x, uv_pverts.permute(0, 2, 3, 1)[..., :2], mode='bilinear', align_corners=False)
    m = uv_pverts[:, -1:]
    uv_delta = uv_delta * m + sd['black_uv_bg'] * (1 - m)
    return torch.cat([uv_gttex, uv_delta, m], dim=1)


def ar_eval_forward(sd, x, vid_c, uvcoords_image, ws, r_list, rendering_kwargs, jitter, u, e4e_results=None,
                    neural_rendering_resolution=128, e4e_training=True, stages=False):
    """inversionNet.AR_eval_forward (return_fake=False), uvnet.py:160-203.
    sd: inversionNet state-dict; x: {'image' [T,3,512,512], 'uv' [T,6,256,256]}; jitter / u pin the renderer's two random
    draws of the T-frame render (evaluation=False inside this call)."""
    from . import triplane as o_tp
    gsd = sub(sd, 'generator')
    T = vid_c.shape[0]
    if ws is None:
        ws = encode(sd, x['image'][0:1], training=e4e_training)
    if e4e_results is None:
        texture_feats = o_sg.synthesis_network(sub(gsd, 'texture_backbone.synthesis'), ws, return_list=True)
        static_feats = o_sg.synthesis_network(sub(gsd, 'backbone.synthesis'), ws, return_list=True)
    else:
        texture_feats, static_feats = e4e_results['texture'], e4e_results['static']
    vid_ws = ws.expand(T, -1, -1)
    y_hat = o_tp.synthesis_with_texture(gsd, vid_ws, [f.expand(T, -1, -1, -1) for f in texture_feats], vid_c, uvcoords_image,
                                        rendering_kwargs, jitter, static_feats=[f.expand(T, -1, -1, -1) for f in static_feats],
                                        evaluation=False, u=u, neural_rendering_resolution=neural_rendering_resolution)
    delta_x = y_hat['image'] - x['image'][:, :3]
    real_vid_uv = get_unet_uvinput(sd, x['uv'], delta_x)
    triplane_input = torch.cat([x['image'][:, :3], delta_x], dim=-3)
    r_list = [None, None] if r_list is None else list(r_list)
    offsets, r_list[0] = texture_unet(sub(sd, 'unet_encoder.texture_unet'), real_vid_uv.unsqueeze(0), r_list[0])
    texture_feats = [f + o for f, o in zip(texture_feats, offsets)] + list(texture_feats[len(offsets):])
    sft, r_list[1] = triplane_unet(sub(sd, 'unet_encoder.triplane_unet'), triplane_input.unsqueeze(0), r_list[1])
    static_feats = o_sg.synthesis_network(sub(gsd, 'backbone.synthesis'), ws, return_list=True, feat_conditions=sft)
    out = {'w': ws, 'texture': texture_feats, 'static': static_feats}
    if stages:
        out.update(e4e_image=y_hat['image'], x_input=real_vid_uv, offsets=offsets, sft=sft)
    return out, r_list
