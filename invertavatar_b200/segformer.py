"""The "improved one-shot" inversion encoder on the B200 engine: Mix-Transformer blocks, the transformer-augmented
UNet decoders and ``uvnet_new.inversionNet`` (reference encoder_inversion/models/{mmseg/mix_transformer,unet_transformer,
uvnet_new}.py; driven by eval_updated_os.py).

Same class names, constructor arguments, sub-module layout and therefore state-dict names as the reference, so
``copy_params_and_buffers(require_all=True)`` (eval_updated_os.py:94) works in both directions.  The torch.nn.Linear / Conv2d /
LayerNorm objects are parameter holders only: no torch arithmetic runs in a forward.  Tokens live as fp32 maps [B,H,W,C] -- an
NHWC image whose pixels are the tokens -- so the reference's flatten / transpose / reshape / permute shuffles do not exist and
every nn.Linear is a 1x1 tensor-core convolution.  A Block is the chain

    ia_layer_norm -> ia_conv_tc (q | kv) -> ia_attention -> ia_conv_tc (proj) -> ia_enc_affine_act (+bias, +residual)
    ia_layer_norm -> ia_conv_tc (fc1) -> ia_dwconv_gelu -> ia_conv_tc (fc2) -> ia_enc_affine_act (+bias, +residual)

and a strided patch embedding is ia_enc_im2col + one GEMM.  ``timm`` (a reference dependency this image lacks) is not needed:
DropPath is the identity at inference, to_2tuple / trunc_normal_ are constructor helpers.  Forward only."""
import math
from functools import partial

import numpy as np
import torch
from torch import nn

from . import runtime as rt
from .encoder import (ConvGRU, DoubleConv, Encoder4Editing, _conv_bias_act, _face_pool, _make_trunk, _nchw, _nhwc, _run_trunk, _sft_head, _sft_heads,
                      inversionNet as _inversionNet_base)


def _init_like_reference(m):
    """The distributions of mix_transformer.py:32-45 (truncated normal 0.02 for linear layers, fan-out normal for convolutions,
    unit LayerNorms), applied once."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


def _tokens(x, H, W):
    """[B,N,C] token tensor -> the engine's [B,H,W,C] map (a view)."""
    B, N, Cc = x.shape
    assert N == H * W
    return x.reshape(B, H, W, Cc)


# ---- mmseg/mix_transformer.py -----------------------------------------------------------------------------------
class DWConv(nn.Module):
    """mix_transformer.py:379-390 (parameter holder; evaluated fused with the GELU that follows it in Mlp)."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def forward(self, x, H, W):
        raise NotImplementedError('DWConv runs fused inside Mlp (ia_dwconv_gelu); call Mlp.forward')


class Mlp(nn.Module):
    """Mix-FFN, mix_transformer.py:18-53."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError('Mlp: only the exact GELU of the reference configuration is fused into ia_dwconv_gelu')
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def run_raw(self, a):
        """a: Split of the normalised tokens -> raw fc2 accumulators [B,H,W,out] (fc2.bias not added)."""
        f = rt.enc_gemm(a, rt.linear_pack(self.fc1, self.fc1.weight))
        g = rt.dwconv_gelu(f, self.fc1.bias, self.dwconv.dwconv)
        return rt.enc_gemm(g, rt.linear_pack(self.fc2, self.fc2.weight))

    def forward(self, x, H, W):
        xm = _tokens(x.float(), H, W).contiguous()
        a, _ = rt.enc_prep([xm])
        y = rt.enc_affine_act(self.run_raw(a), shift=self.fc2.bias)
        return y.reshape(x.shape[0], H * W, -1)


class Attention(nn.Module):
    """Efficient self-attention, mix_transformer.py:56-115."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.sr_ratio = sr_ratio
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)

    def run_raw(self, a, x32):
        """a: Split of the normalised tokens [B,H,W,C]; x32: the same tokens in fp32 (needed when sr_ratio > 1, else None)
        -> raw proj accumulators [B,H,W,C] (proj.bias not added)."""
        if self.sr_ratio == 1 and rt.attention_tc_available(self.dim // self.num_heads, self.q.bias, self.kv.bias):
            # tensor-core attention: the q / kv projections emit their bf16 hi/lo operands straight from the GEMM epilogue
            q = rt.enc_gemm(a, rt.linear_pack(self.q, self.q.weight), emit_split=True)
            kv = rt.enc_gemm(a, rt.linear_pack(self.kv, self.kv.weight), emit_split=True)
            att = rt.attention_tc(q, kv, self.num_heads, self.scale)
            return rt.enc_gemm(att, rt.linear_pack(self.proj, self.proj.weight))
        q = rt.enc_gemm(a, rt.linear_pack(self.q, self.q.weight))
        if self.sr_ratio > 1:
            pk = rt.im2col_pack(self.sr)
            s = rt.enc_gemm(rt.enc_im2col([x32], self.sr_ratio, self.sr_ratio, 0, K_pad=pk.Cin_pad), pk)
            a_kv, _ = rt.layer_norm(s, self.norm, pre_bias=self.sr.bias)
        else:
            a_kv = a
        kv = rt.enc_gemm(a_kv, rt.linear_pack(self.kv, self.kv.weight))
        att = rt.attention(q, kv, self.num_heads, self.scale, q_bias=self.q.bias, kv_bias=self.kv.bias)
        return rt.enc_gemm(att, rt.linear_pack(self.proj, self.proj.weight))

    def forward(self, x, H, W):
        xm = _tokens(x.float(), H, W).contiguous()
        a, _ = rt.enc_prep([xm])
        y = rt.enc_affine_act(self.run_raw(a, xm), shift=self.proj.bias)
        return y.reshape(x.shape[0], H * W, -1)


class Block(nn.Module):
    """mix_transformer.py:118-156.  ``drop_path`` is kept as a constructor argument (stochastic depth acts in training only)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop, sr_ratio=sr_ratio)
        self.drop_path = nn.Identity()
        self.drop_path_prob = float(drop_path)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def run_nhwc(self, x):
        """x [B,H,W,C] fp32 tokens -> same shape."""
        if self.training and self.drop_path_prob > 0:
            raise NotImplementedError('Block: stochastic depth (training) is not part of the inference path')
        sr = self.attn.sr_ratio > 1
        a, x32 = rt.layer_norm(x, self.norm1, want32=sr)
        x = rt.enc_affine_act(self.attn.run_raw(a, x32), shift=self.attn.proj.bias, res=x)
        a, _ = rt.layer_norm(x, self.norm2)
        return rt.enc_affine_act(self.mlp.run_raw(a), shift=self.mlp.fc2.bias, res=x)

    def forward(self, x, H, W):
        y = self.run_nhwc(_tokens(x.float(), H, W).contiguous())
        return y.reshape(x.shape[0], H * W, -1)


class OverlapPatchEmbed(nn.Module):
    """mix_transformer.py:159-198."""

    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = img_size if isinstance(img_size, tuple) else (img_size, img_size)
        patch_size = patch_size if isinstance(patch_size, tuple) else (patch_size, patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.H, self.W = img_size[0] // patch_size[0], img_size[1] // patch_size[1]
        self.num_patches = self.H * self.W
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride, padding=(patch_size[0] // 2, patch_size[1] // 2))
        self.norm = nn.LayerNorm(embed_dim)

    def run_nhwc(self, srcs):
        """srcs: [B,H,W,Ci] views (tensor or (tensor, pixel_shuffle)) concatenated along channels -> fp32 tokens [B,h,w,embed_dim]."""
        pk = rt.im2col_pack(self.proj)
        a = rt.enc_im2col(srcs, self.patch_size[0], self.proj.stride[0], self.proj.padding[0], K_pad=pk.Cin_pad)
        _, x = rt.layer_norm(rt.enc_gemm(a, pk), self.norm, pre_bias=self.proj.bias, want_split=False, want32=True)
        return x

    def forward(self, x):
        t = self.run_nhwc([_nhwc(x.float())])
        B, H, W, Cc = t.shape
        return t.reshape(B, H * W, Cc), H, W


class MixVisionTransformer(nn.Module):
    """mix_transformer.py:201-376 (forward_features; the classification head is commented out in the reference)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dims=[64, 128, 256, 512], num_heads=[1, 2, 4, 8],
                 mlp_ratios=[4, 4, 4, 4], qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, depths=[3, 4, 6, 3], sr_ratios=[8, 4, 2, 1]):
        super().__init__()
        self.num_classes = num_classes
        self.depths = depths
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for i in range(4):
            setattr(self, f'patch_embed{i + 1}', OverlapPatchEmbed(img_size=img_size // (1 if i == 0 else 2 ** (i + 1)), patch_size=7 if i == 0 else 3,
                                                                   stride=4 if i == 0 else 2, in_chans=in_chans if i == 0 else embed_dims[i - 1],
                                                                   embed_dim=embed_dims[i]))
        for i in range(4):
            setattr(self, f'block{i + 1}', nn.ModuleList([
                Block(dim=embed_dims[i], num_heads=num_heads[i], mlp_ratio=mlp_ratios[i], qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                      attn_drop=attn_drop_rate, drop_path=dpr[cur + j], norm_layer=norm_layer, sr_ratio=sr_ratios[i]) for j in range(depths[i])]))
            setattr(self, f'norm{i + 1}', norm_layer(embed_dims[i]))
            cur += depths[i]
        self.apply(_init_like_reference)

    def forward_features(self, x):
        outs = []
        srcs = [_nhwc(x.float())]
        for i in range(4):
            t = getattr(self, f'patch_embed{i + 1}').run_nhwc(srcs)
            for blk in getattr(self, f'block{i + 1}'):
                t = blk.run_nhwc(t)
            _, t = rt.layer_norm(t, getattr(self, f'norm{i + 1}'), want_split=False, want32=True)
            outs.append(_nchw(t))
            srcs = [t]
        return outs

    def forward(self, x):
        return self.forward_features(x)


class MLP(nn.Module):
    """Linear embedding, mix_transformer.py:393-404: [B,C,H,W] -> tokens [B,HW,embed_dim]."""

    def __init__(self, input_dim=2048, embed_dim=768):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)

    def forward(self, x):
        a, _ = rt.enc_prep([_nhwc(x.float())])
        y = rt.enc_affine_act(rt.enc_gemm(a, rt.linear_pack(self.proj, self.proj.weight)), shift=self.proj.bias)
        return y.reshape(x.shape[0], -1, y.shape[-1])


class transformer_block(nn.Module):
    """mix_transformer.py:453-472: 7x7 stride-2 patch embedding to ``embed_dim`` channels, ``num_vit`` Blocks (4 heads, mlp_ratio 2),
    LayerNorm, PixelShuffle(2) back to the input resolution, 1x1 convolution back to ``in_chans``."""

    def __init__(self, in_chans=256, embed_dim=1024, num_vit=2):
        super().__init__()
        self.patch_embed = OverlapPatchEmbed(img_size=0, stride=2, in_chans=in_chans, embed_dim=embed_dim)
        self.ViT = nn.ModuleList([Block(dim=embed_dim, num_heads=4, mlp_ratio=2, sr_ratio=1) for _ in range(num_vit)])
        self.pixel_shuffle = nn.PixelShuffle(upscale_factor=2)
        self.mlp = nn.Conv2d(embed_dim // 4, in_chans, kernel_size=1)
        self.norm = nn.LayerNorm(embed_dim)
        self.apply(_init_like_reference)

    def run_nhwc(self, srcs):
        """srcs: [B,H,W,Ci] views (sum Ci = in_chans) -> [B,H,W,in_chans] fp32."""
        t = self.patch_embed.run_nhwc(srcs)
        for blk in self.ViT:
            t = blk.run_nhwc(t)
        _, t = rt.layer_norm(t, self.norm, want_split=False, want32=True)
        return _conv_bias_act([(t, self.pixel_shuffle.upscale_factor)], self.mlp)

    def forward(self, f):
        return _nchw(self.run_nhwc([_nhwc(f.float())]))


# ---- unet_transformer.py ----------------------------------------------------------------------------------------
class UpLayer(nn.Module):
    """unet_transformer.py:523-547."""

    def __init__(self, in_channels, out_channels, upscale_factor=2, use_gru=False, num_vit=0):
        super().__init__()
        self.up = nn.PixelShuffle(upscale_factor=upscale_factor)
        self.conv = DoubleConv(in_channels, out_channels)
        self.conv_gru = ConvGRU(out_channels, out_act_prelu=False) if use_gru else None
        self.use_vit = num_vit > 0
        self.transformer = transformer_block(in_chans=in_channels, num_vit=num_vit) if self.use_vit else None

    def run_nhwc(self, x1, x2=None, T=0, r=None):
        srcs = [(x1, self.up.upscale_factor)] if x2 is None else [x2, (x1, self.up.upscale_factor)]
        if self.use_vit:
            srcs = [self.transformer.run_nhwc(srcs)]
        x = self.conv.run_nhwc(srcs)
        if self.conv_gru is None:
            return x
        return self.conv_gru.run_nhwc(x.unflatten(0, (-1, T)), r, False)

    def forward(self, x1, x2=None, T=0, r=None):
        o = self.run_nhwc(_nhwc(x1), None if x2 is None else _nhwc(x2), T, None if r is None else rt.to_nhwc(r))
        if self.conv_gru is None:
            return _nchw(o)
        return _nchw(o[0]), _nchw(o[1])


class _SegformerDecoderBase(nn.Module):
    def _build(self, inp_ch, res, use_gru, num_vits):
        self.res = res
        self.use_gru = use_gru
        self.face_pool = None if res is None else torch.nn.AdaptiveAvgPool2d((res, res))
        _make_trunk(self, inp_ch)
        self.up1 = UpLayer(1024, 512, upscale_factor=1, use_gru=use_gru, num_vit=num_vits[0])
        self.up2 = UpLayer(384, 384, use_gru=use_gru, num_vit=num_vits[1])
        self.up3 = UpLayer(224, 256, use_gru=use_gru, num_vit=num_vits[2])
        self.up4 = UpLayer(128, 96, use_gru=use_gru, num_vit=num_vits[3])

    def _trunk_decoder(self, x, r_list):
        if x.dim() == 5:
            T = x.shape[1]
            x = x.flatten(0, 1)
        else:
            T = 1
        xn = _nhwc(x.float())
        if self.face_pool is not None:
            xn = _face_pool(xn, self.res)
        x, f = _run_trunk(self, xn, taps=(2, 6, 20, 21))
        c0, c1, c2, c3 = f[2], f[6], f[20], f[21]
        if not self.use_gru:
            t1 = self.up1.run_nhwc(x, c3)
            t2 = self.up2.run_nhwc(t1, c2)
            t3 = self.up3.run_nhwc(t2, c1)
            t4 = self.up4.run_nhwc(t3, c0)
            return (t1, t2, t3, t4), None
        from .encoder import _repeat_T, _state_in
        r_list = [None] * 4 if r_list is None else list(r_list)
        t1, r0 = self.up1.run_nhwc(x, c3, T, _state_in(r_list[0]))
        t2, r1 = self.up2.run_nhwc(_repeat_T(t1, T), c2, T, _state_in(r_list[1]))
        t3, r2 = self.up3.run_nhwc(_repeat_T(t2, T), c1, T, _state_in(r_list[2]))
        t4, r3 = self.up4.run_nhwc(_repeat_T(t3, T), c0, T, _state_in(r_list[3]))
        return (t1, t2, t3, t4), [_nchw(r) for r in (r0, r1, r2, r3)]


class TriPlanefeat_SegformerDecoder(_SegformerDecoderBase):
    """Texture decoder, unet_transformer.py:255-337."""

    def __init__(self, inp_ch, sft_half=True, res=None, use_gru=False):
        super().__init__()
        self.sft_half = sft_half
        self._build(inp_ch, res, use_gru, (4, 4, 3, 3))
        self.outconv0 = nn.Conv2d(384, 32, kernel_size=1, padding=0)
        self.outconv1 = nn.Conv2d(384, 512, kernel_size=1, padding=0)
        self.outconv2 = nn.Conv2d(256, 512, kernel_size=1, padding=0)
        self.outconv3 = nn.Conv2d(96, 256, kernel_size=1, padding=0)

    def forward(self, x, r_list=None, return_list=True):
        (t1, t2, t3, t4), r_list = self._trunk_decoder(x, r_list)
        out_list = [_nchw(_conv_bias_act([t], conv)) for t, conv in ((t2, self.outconv0), (t2, self.outconv1), (t3, self.outconv2),
                                                                    (t4, self.outconv3))]
        return (out_list, r_list) if self.use_gru else out_list


class TriPlaneSFTfeat_SegformerDecoder(_SegformerDecoderBase):
    """Tri-plane SFT decoder, unet_transformer.py:340-450."""

    def __init__(self, inp_ch, sft_half=True, res=None, use_gru=False):
        super().__init__()
        self.sft_half = sft_half
        self._build(inp_ch, res, use_gru, (4, 4, 3, 2))
        self.head = nn.PixelShuffle(upscale_factor=2)
        self.final_head = nn.Sequential(nn.Conv2d(24, 96, kernel_size=3, padding=1), nn.PReLU(96),
                                        nn.Conv2d(96, 96, kernel_size=3, padding=1), nn.PReLU(96))
        self.block_resolutions = [2 ** i for i in range(int(np.log2(16)), int(np.log2(256)) + 1)]
        channels_dict = {res: min(32768 // res, 512) for res in self.block_resolutions}
        body_outchannels_dict = {16: 512, 32: 384, 64: 256, 128: 96, 256: 96}
        for res in self.block_resolutions:
            out_channels = body_outchannels_dict[res]
            sft_out_channels = channels_dict[res] // 2 if self.sft_half else channels_dict[res]
            for kind in ('scale', 'shift'):
                setattr(self, f'condition_{kind}{res}', nn.Sequential(
                    nn.Conv2d(out_channels, out_channels, 3, 1, 1), nn.LeakyReLU(0.2, True),
                    nn.Conv2d(out_channels, sft_out_channels, 3, 1, 1)))

    def _head(self, res, t):
        return _sft_head(self, res, t)

    def forward(self, x, r_list=None):
        (t1, t2, t3, t4), r_list = self._trunk_decoder(x, r_list)
        out = _sft_heads(self, (t1, t2, t3, t4))
        return (out, r_list) if self.use_gru else out


# ---- uvnet_new.py -----------------------------------------------------------------------------------------------
class improved_os_unet_encoder(nn.Module):
    """uvnet_new.py:13-21."""

    def __init__(self, encoding_texture=False, encoding_triplane=False):
        super().__init__()
        self.texture_unet = TriPlanefeat_SegformerDecoder(inp_ch=7, res=256) if encoding_texture else None
        self.triplane_unet = TriPlaneSFTfeat_SegformerDecoder(inp_ch=6, res=256) if encoding_triplane else None

    def forward(self, x):
        raise NotImplementedError


class inversionNet(_inversionNet_base):
    """uvnet_new.py:24-162: the one-shot encoder of eval_updated_os.py -- e4e + the two SegFormer-style decoders (no ConvGRU, no
    AR_eval_forward).  ``encode`` / ``get_unet_uvinput`` / ``forward`` are the implementations shared with uvnet.inversionNet
    (the reference's two files restate them verbatim, uvnet_new.py:107-157 == uvnet.py:107-157)."""

    def __init__(self, G_kwargs=None, generator=None, encoding_texture=True, encoding_triplane=False):
        super().__init__(G_kwargs=G_kwargs, generator=generator, encoding_texture=False, encoding_triplane=False)
        self.unet_encoder = improved_os_unet_encoder(encoding_texture=encoding_texture, encoding_triplane=encoding_triplane)

    def AR_eval_forward(self, *args, **kwargs):
        raise AttributeError('uvnet_new.inversionNet has no AR_eval_forward (the reference class does not define it)')


__all__ = ['DWConv', 'Mlp', 'Attention', 'Block', 'OverlapPatchEmbed', 'MixVisionTransformer', 'MLP', 'transformer_block', 'UpLayer',
           'TriPlanefeat_SegformerDecoder', 'TriPlaneSFTfeat_SegformerDecoder', 'improved_os_unet_encoder', 'inversionNet', 'Encoder4Editing',
           'partial']
