"""CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py) of the "improved one-shot" inversion encoder
(SURVEY 8f-4): the Mix-Transformer blocks, the transformer-augmented UNet decoders and ``uvnet_new.inversionNet.forward``.

Functional restatement over a state-dict (same key names as the reference modules); every function cites the reference
file:line it follows (paths relative to /root/reference/encoder_inversion/models).  Linear / conv / layer_norm / softmax /
gelu / pixel_shuffle are PyTorch ATen in both the reference and here (third-party arithmetic, oracle/__init__.py).

The reference imports three helpers from ``timm`` (DropPath, to_2tuple, trunc_normal_; mmseg/mix_transformer.py:11-13), a
dependency that is absent from this image.  DropPath is the identity outside training and for drop_path == 0 (the only
configuration ``transformer_block`` builds, mix_transformer.py:458), the other two only act in constructors, so the forward
restated here does not depend on timm; tests/golden/make_golden_segformer.py pins it against the unmodified reference run
with a three-function stub of that package.
"""
import torch
import torch.nn.functional as F

from . import encoder as o_enc
from . import stylegan2 as o_sg

LN_EPS_DEFAULT = 1e-5     # torch.nn.LayerNorm default (transformer_block); MixVisionTransformer passes eps=1e-6


def sub(sd, prefix):
    return o_sg.sub(sd, prefix)


# ---- mmseg/mix_transformer.py ------------------------------------------------------------------------------
def layer_norm(sd, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd['weight'], sd['bias'], eps)


def dwconv(sd, x, H, W):
    """DWConv, mix_transformer.py:379-390: tokens -> image -> depthwise 3x3 (+bias) -> tokens."""
    B, N, Cc = x.shape
    y = x.transpose(1, 2).reshape(B, Cc, H, W)
    y = F.conv2d(y, sd['dwconv.weight'], sd['dwconv.bias'], padding=1, groups=Cc)
    return y.flatten(2).transpose(1, 2)


def mix_ffn(sd, x, H, W):
    """Mlp, mix_transformer.py:18-53: fc1 -> depthwise 3x3 -> GELU (exact, erf) -> fc2 (dropout p=0)."""
    x = F.linear(x, sd['fc1.weight'], sd['fc1.bias'])
    x = F.gelu(dwconv(sub(sd, 'dwconv'), x, H, W))
    return F.linear(x, sd['fc2.weight'], sd['fc2.bias'])


def attention(sd, x, H, W, num_heads, sr_ratio, eps):
    """Attention, mix_transformer.py:56-115: multi-head attention whose keys / values come from a sr_ratio-strided
    patchified + layer-normed copy of the tokens when sr_ratio > 1."""
    B, N, Cc = x.shape
    hd = Cc // num_heads
    scale = hd ** -0.5
    q = F.linear(x, sd['q.weight'], sd.get('q.bias')).reshape(B, N, num_heads, hd).permute(0, 2, 1, 3)
    if sr_ratio > 1:
        x_ = x.permute(0, 2, 1).reshape(B, Cc, H, W)
        x_ = F.conv2d(x_, sd['sr.weight'], sd['sr.bias'], stride=sr_ratio).reshape(B, Cc, -1).permute(0, 2, 1)
        x_ = layer_norm(sub(sd, 'norm'), x_, LN_EPS_DEFAULT)      # Attention builds its own nn.LayerNorm(dim): default eps
    else:
        x_ = x
    kv = F.linear(x_, sd['kv.weight'], sd.get('kv.bias')).reshape(B, -1, 2, num_heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, Cc)
    return F.linear(y, sd['proj.weight'], sd['proj.bias'])


def block(sd, x, H, W, num_heads, sr_ratio, eps):
    """Block, mix_transformer.py:118-156 (DropPath = identity at inference)."""
    x = x + attention(sub(sd, 'attn'), layer_norm(sub(sd, 'norm1'), x, eps), H, W, num_heads, sr_ratio, eps)
    return x + mix_ffn(sub(sd, 'mlp'), layer_norm(sub(sd, 'norm2'), x, eps), H, W)


def overlap_patch_embed(sd, x, patch, stride):
    """OverlapPatchEmbed, mix_transformer.py:159-198: conv(patch, stride, pad patch//2) -> tokens -> LayerNorm (default eps)."""
    x = F.conv2d(x, sd['proj.weight'], sd['proj.bias'], stride=stride, padding=patch // 2)
    H, W = x.shape[2:]
    return layer_norm(sub(sd, 'norm'), x.flatten(2).transpose(1, 2), LN_EPS_DEFAULT), H, W


def mix_vision_transformer(sd, x, depths, num_heads=(1, 2, 5, 8), sr_ratios=(8, 4, 2, 1), eps=1e-6):
    """MixVisionTransformer.forward_features, mix_transformer.py:332-370 -> the four stage outputs [B,C_i,H_i,W_i]."""
    B = x.shape[0]
    outs = []
    for i in range(4):
        x, H, W = overlap_patch_embed(sub(sd, f'patch_embed{i + 1}'), x, 7 if i == 0 else 3, 4 if i == 0 else 2)
        for j in range(depths[i]):
            x = block(sub(sd, f'block{i + 1}.{j}'), x, H, W, num_heads[i], sr_ratios[i], eps)
        x = layer_norm(sub(sd, f'norm{i + 1}'), x, eps)
        x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def transformer_block(sd, f, num_vit):
    """transformer_block, mix_transformer.py:453-472: 7x7 stride-2 patch embedding to 1024 channels, num_vit Blocks (4 heads,
    mlp_ratio 2, no spatial reduction, no q/kv bias), LayerNorm, PixelShuffle(2) back to the input resolution, 1x1 convolution
    back to the input channel count."""
    B = f.shape[0]
    x, H, W = overlap_patch_embed(sub(sd, 'patch_embed'), f, 7, 2)
    for j in range(num_vit):
        x = block(sub(sd, f'ViT.{j}'), x, H, W, 4, 1, LN_EPS_DEFAULT)
    x = layer_norm(sub(sd, 'norm'), x, LN_EPS_DEFAULT)
    x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2)
    return F.conv2d(F.pixel_shuffle(x, 2), sd['mlp.weight'], sd['mlp.bias'])


# ---- unet_transformer.py -----------------------------------------------------------------------------------
def up_layer(sd, x1, x2, upscale, num_vit, training):
    """UpLayer.forward (use_gru=False), unet_transformer.py:523-547: PixelShuffle -> concat -> transformer_block -> DoubleConv."""
    if upscale > 1:
        x1 = F.pixel_shuffle(x1, upscale)
    x = x1 if x2 is None else torch.cat([x2, x1], dim=1)
    if num_vit > 0:
        x = transformer_block(sub(sd, 'transformer'), x, num_vit)
    return o_enc.double_conv(sub(sd, 'conv'), x, training)


def _trunk_and_decoder(sd, x, num_vits, bn_trunk_training, bn_decoder_training, res=256):
    """Shared part of the two decoders (unet_transformer.py:283-296,326-337 / 391-404,436-450)."""
    if x.dim() == 5:
        x = x.flatten(0, 1)
    x = o_enc.face_pool(x, res)
    x, f = o_enc.ir_se50_trunk(sd, x, taps=(2, 6, 20, 21), training=bn_trunk_training)
    c0, c1, c2, c3 = f[2], f[6], f[20], f[21]
    t1 = up_layer(sub(sd, 'up1'), x, c3, 1, num_vits[0], bn_decoder_training)
    t2 = up_layer(sub(sd, 'up2'), t1, c2, 2, num_vits[1], bn_decoder_training)
    t3 = up_layer(sub(sd, 'up3'), t2, c1, 2, num_vits[2], bn_decoder_training)
    t4 = up_layer(sub(sd, 'up4'), t3, c0, 2, num_vits[3], bn_decoder_training)
    return t1, t2, t3, t4


def texture_segformer_decoder(sd, x, bn_trunk_training=False, bn_decoder_training=False):
    """TriPlanefeat_SegformerDecoder.forward (use_gru=False), unet_transformer.py:283-337 -> [4 offsets]."""
    t1, t2, t3, t4 = _trunk_and_decoder(sd, x, (4, 4, 3, 3), bn_trunk_training, bn_decoder_training)
    return [F.conv2d(t2, sd['outconv0.weight'], sd['outconv0.bias']), F.conv2d(t2, sd['outconv1.weight'], sd['outconv1.bias']),
            F.conv2d(t3, sd['outconv2.weight'], sd['outconv2.bias']), F.conv2d(t4, sd['outconv3.weight'], sd['outconv3.bias'])]


def triplane_segformer_decoder(sd, x, bn_trunk_training=False, bn_decoder_training=False):
    """TriPlaneSFTfeat_SegformerDecoder.forward (use_gru=False), unet_transformer.py:391-450 -> {res: stack(scale, shift)}."""
    t1, t2, t3, t4 = _trunk_and_decoder(sd, x, (4, 4, 3, 2), bn_trunk_training, bn_decoder_training)
    t5 = F.pixel_shuffle(t4, 2)
    t5 = o_enc.prelu(F.conv2d(t5, sd['final_head.0.weight'], sd['final_head.0.bias'], padding=1), sd['final_head.1.weight'])
    t5 = o_enc.prelu(F.conv2d(t5, sd['final_head.2.weight'], sd['final_head.2.bias'], padding=1), sd['final_head.3.weight'])
    out = {}
    for res, t in zip((16, 32, 64, 128, 256), (t1, t2, t3, t4, t5)):
        out[res] = torch.stack([o_enc._sft_head(sd, f'condition_scale{res}', t), o_enc._sft_head(sd, f'condition_shift{res}', t)])
    return out


# ---- uvnet_new.py ------------------------------------------------------------------------------------------
def forward(sd, x, cam, uvcoords_image, rendering_kwargs, draws, e4e_results=None, training=False, neural_rendering_resolution=None):
    """uvnet_new.inversionNet.forward, uvnet_new.py:123-157.  sd: inversionNet state-dict; x: {'image' [B,3,S,S], 'uv' [B,6,256,256]};
    draws: [(jitter, u), (jitter, u)] for the two synthesis_withTexture calls (evaluation=False: random-u importance sampling).
    ``training``: the BatchNorm mode of the whole encoder (eval_updated_os.py:93 builds it with .eval())."""
    from . import triplane as o_tp
    g_sd = sub(sd, 'generator')
    if e4e_results is None:
        ws = o_enc.encode(sd, x['image'][:, :3], n_styles=_n_styles(sd), training=training)
        tex = o_sg.synthesis_network(sub(g_sd, 'texture_backbone.synthesis'), ws, return_list=True)
        sta = o_sg.synthesis_network(sub(g_sd, 'backbone.synthesis'), ws, return_list=True)
    else:
        ws, tex, sta = e4e_results['w'], e4e_results['texture'], e4e_results['static']
    y_hat = o_tp.synthesis_with_texture(g_sd, ws, tex, cam, uvcoords_image, rendering_kwargs, draws[0][0], static_feats=sta,
                                        evaluation=False, u=draws[0][1], neural_rendering_resolution=neural_rendering_resolution)
    delta_x = y_hat['image'] - x['image'][:, :3]
    x_input = o_enc.get_unet_uvinput(sd, x['uv'], delta_x)              # uvnet_new.py:117-121 == uvnet.py:117-121
    offsets = texture_segformer_decoder(sub(sd, 'unet_encoder.texture_unet'), x_input, training, training)
    texture_feats = [f + o for f, o in zip(tex, offsets)] + list(tex[len(offsets):])
    sft = triplane_segformer_decoder(sub(sd, 'unet_encoder.triplane_unet'), torch.cat([x['image'][:, :3], delta_x], dim=1), training, training)
    static_feats = o_sg.synthesis_network(sub(g_sd, 'backbone.synthesis'), ws, return_list=True, feat_conditions=sft)
    out = o_tp.synthesis_with_texture(g_sd, ws, texture_feats, cam, uvcoords_image, rendering_kwargs, draws[1][0], static_feats=static_feats,
                                      evaluation=False, u=draws[1][1], neural_rendering_resolution=neural_rendering_resolution)
    out['texture'], out['static'], out['w'], out['e4e_image'], out['x_input'] = texture_feats, static_feats, ws, y_hat['image'], x_input
    return out


def _n_styles(sd):
    return 1 + max(int(k.split('.')[2]) for k in sd if k.startswith('encoder.styles.'))
