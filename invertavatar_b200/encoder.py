"""Inversion encoder on the B200 engine: e4e (IR-SE50 + FPN + GradualStyleBlocks), the texture UNet and the tri-plane
SFT UNet with their DoubleConv / ConvGRU decoders, and ``inversionNet`` (reference encoder_inversion/models/
{helpers,e4e,unet_encoders,uvnet}.py).

Same classes, constructor arguments, sub-module layout and therefore state-dict names as the reference, so
``copy_params_and_buffers(require_all=True)`` works in both directions (eval_seq.py:94) and parameters drawn under
``torch.manual_seed`` match the reference's.  The torch.nn.Conv2d / BatchNorm2d / PReLU objects are parameter holders
only: no torch arithmetic runs in a forward.  A forward is a chain of

    ia_enc_prep (concat / PixelShuffle / stride views + BatchNorm affine + PReLU -> bf16 hi/lo operand)
    -> ia_conv_tc (tcgen05 implicit GEMM, 3-term split, fp32 accumulators)
    -> ia_enc_affine_act (bias / BatchNorm / PReLU / SE gate / shortcut add)

plus the small normalisation, pooling and ConvGRU gating kernels of csrc/ia_encoder.cu.  Activations are fp32
channels-last between layers; tensors handed to the caller are logical NCHW views of those buffers.  BatchNorm follows
each module's ``training`` flag exactly like torch (eval_seq.py:92-97 leaves e4e and the UNet decoders in train mode:
batch statistics).  Forward only."""
import os
from collections import namedtuple

import numpy as np
import torch
from torch import nn
from torch.nn import BatchNorm2d, Conv2d, Module, PReLU, Sequential

from . import runtime as rt
from .stylegan2 import FullyConnectedLayer


# ---- helpers.py ------------------------------------------------------------------------------------------------
class Bottleneck(namedtuple('Block', ['in_channel', 'depth', 'stride'])):
    """A named tuple describing a ResNet block (helpers.py:26-27)."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    """helpers.py:34-59."""
    units = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in units:
        raise ValueError('Invalid number of layers: {}. Must be one of [50, 100, 152]'.format(num_layers))
    u = units[num_layers]
    return [get_block(64, 64, u[0]), get_block(64, 128, u[1]), get_block(128, 256, u[2]), get_block(256, 512, u[3])]


def _nhwc(x):
    """logical NCHW (any strides) -> [B,H,W,C]-indexed view (no copy)."""
    return x.permute(0, 2, 3, 1)


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _sub(x, s):
    return x if s == 1 else x[:, ::s, ::s]


def _conv_bias_act(x_srcs, conv, slope=None, lrelu=1.0, in_scale=None, in_shift=None, in_slope=None, in_lrelu=1.0, stride=1,
                   slope2=None, res=None, out=None):
    """conv(+bias) of cat(x_srcs) with optional input-side affine/activation and output-side PReLU / leaky ReLU."""
    pack = rt.ConvPack.current(conv, '_ia_pack', conv.weight, need_wsq=False)
    a, _ = rt.enc_prep(x_srcs, scale=in_scale, shift=in_shift, slope=in_slope, lrelu=in_lrelu, C_pad=pack.Cin_pad)
    raw = _sub(rt.enc_conv(a, conv), stride)
    if lrelu != 1.0:
        return rt.enc_affine_act(raw, shift=conv.bias, act='lrelu', alpha=lrelu, res=res, out=out)
    return rt.enc_affine_act(raw, shift=conv.bias, slope1=slope, slope2=slope2, res=res, out=out)


class SEModule(Module):
    """helpers.py:62-80."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def gate(self, x, scale=None, shift=None):
        """sigmoid(fc2(relu(fc1(mean_hw(x*scale+shift))))) -> [B,C]."""
        Cr, Cc = self.fc1.weight.shape[:2]
        if rt.enc_epilogue_fusion():      # pool + fc1 + ReLU + fc2 + sigmoid in one launch
            return rt.enc_se_gate(x, scale, shift, self.fc1.weight.reshape(Cr, Cc), self.fc2.weight.reshape(Cc, Cr))
        pooled = rt.enc_global_pool(x, scale, shift)
        g = rt.fully_connected(pooled, self.fc1.weight.reshape(Cr, Cc), None, act='relu', act_gain=1.0)
        return rt.fully_connected(g, self.fc2.weight.reshape(Cc, Cr), None, act='sigmoid', act_gain=1.0)

    def forward(self, x):
        xn = _nhwc(x)
        return _nchw(rt.enc_affine_act(xn, gate=self.gate(xn)))


class bottleneck_IR_SE(Module):
    """helpers.py:102-124."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        self.stride = stride
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = Sequential(Conv2d(in_channel, depth, (1, 1), stride, bias=False), BatchNorm2d(depth))
        self.res_layer = Sequential(
            BatchNorm2d(in_channel),
            Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
            PReLU(depth),
            Conv2d(depth, depth, (3, 3), stride, 1, bias=False),
            BatchNorm2d(depth),
            SEModule(depth, 16))

    def opening_affine(self):
        """(scale, shift, C_pad) of the BatchNorm that opens this unit when it is a fixed affine (eval mode), else None: the unit in
        front can then emit this unit's first operand from its closing pass (ia_enc_affine_params.e_*)."""
        bn0, conv1 = self.res_layer[0], self.res_layer[1]
        if bn0.training or not bn0.track_running_stats:
            return None
        sc0, sh0 = rt.enc_bn_fold(bn0, [])
        return sc0, sh0, rt.ConvPack.current(conv1, '_ia_pack', conv1.weight, need_wsq=False).Cin_pad

    def run_nhwc(self, x, a=None, emit=None):
        """x [B,H,W,Cin] fp32 -> [B,H/s,W/s,depth].  a: this unit's first operand split(bn0(x)) when the caller already has it;
        emit: the next unit's opening_affine() -> returns (y, split(bn0_next(y)))."""
        s = self.stride
        bn0, conv1, prelu, conv2, bn4, se = self.res_layer
        if a is None:
            sc0, sh0 = rt.enc_bn_fold(bn0, [x])
            a, _ = rt.enc_prep([x], scale=sc0, shift=sh0, C_pad=rt.ConvPack.current(conv1, '_ia_pack', conv1.weight, need_wsq=False).Cin_pad)
        pad2 = rt.ConvPack.current(conv2, '_ia_pack', conv2.weight, need_wsq=False).Cin_pad
        if rt.enc_epilogue_fusion():
            a = rt.enc_conv_act(a, conv1, pad2, slope=prelu.weight)      # PReLU + operand emission inside conv1's epilogue
        else:
            raw1 = rt.enc_conv(a, conv1)
            a, _ = rt.enc_prep([raw1], slope=prelu.weight, C_pad=pad2)
        raw2 = _sub(rt.enc_conv(a, conv2), s)      # stride-s convolution = the stride-1 result sampled every s pixels
        sc4, sh4 = rt.enc_bn_fold(bn4, [raw2])
        gate = se.gate(raw2, sc4, sh4)
        xs = _sub(x, s)                            # MaxPool2d(1, s) / the input of the strided 1x1 shortcut convolution
        if isinstance(self.shortcut_layer, Sequential):
            conv_s, bn_s = self.shortcut_layer
            a, _ = rt.enc_prep([xs], C_pad=rt.ConvPack.current(conv_s, '_ia_pack', conv_s.weight, need_wsq=False).Cin_pad)
            raw_s = rt.enc_conv(a, conv_s, alg_stride=1)      # operand already sub-sampled
            rs, rsh = rt.enc_bn_fold(bn_s, [raw_s])
            return rt.enc_affine_act(raw2, scale=sc4, shift=sh4, gate=gate, res=raw_s, res_scale=rs, res_shift=rsh, emit=emit)
        return rt.enc_affine_act(raw2, scale=sc4, shift=sh4, gate=gate, res=xs, emit=emit)

    def forward(self, x):
        return _nchw(self.run_nhwc(_nhwc(x)))


def _make_trunk(module, inp_ch):
    """input_layer + body of IR-SE50 (e4e.py:72-82, unet_encoders.py:110-119)."""
    module.input_layer = Sequential(Conv2d(inp_ch, 64, (3, 3), 1, 1, bias=False), BatchNorm2d(64), PReLU(64))
    modules = []
    for block in get_blocks(num_layers=50):
        for b in block:
            modules.append(bottleneck_IR_SE(b.in_channel, b.depth, b.stride))
    module.body = Sequential(*modules)


def _run_trunk(module, x, taps):
    """x: [B,H,W,C]-indexed view -> (final activation, {tap index: activation}), all fp32 NHWC."""
    conv, bn, prelu = module.input_layer
    a, _ = rt.enc_prep([x], C_pad=rt.ConvPack.current(conv, '_ia_pack', conv.weight, need_wsq=False).Cin_pad)
    raw = rt.enc_conv(a, conv)
    sc, sh = rt.enc_bn_fold(bn, [raw])
    x = rt.enc_affine_act(raw, scale=sc, shift=sh, slope1=prelu.weight)
    feats = {}
    blocks = list(module.body)
    a = None
    for i, blk in enumerate(blocks):
        # a unit whose successor opens with an eval-mode BatchNorm also writes the successor's first operand
        nxt = blocks[i + 1].opening_affine() if i + 1 < len(blocks) else None
        out = blk.run_nhwc(x, a=a, emit=nxt)
        x, a = out if nxt is not None else (out, None)
        if i in taps:
            feats[i] = x
    return x, feats


def _face_pool(x_nhwc_view, res):
    """AdaptiveAvgPool2d((res,res)) when the input is not already res wide (integer ratios only on this path)."""
    W = x_nhwc_view.shape[2]
    if W == res:
        return x_nhwc_view
    if W % res != 0 or x_nhwc_view.shape[1] % res != 0:
        raise RuntimeError(f'face_pool: only integer down-sampling ratios are supported (got {W} -> {res})')
    return rt.enc_avgpool(x_nhwc_view, W // res)


# ---- e4e.py ----------------------------------------------------------------------------------------------------
class GradualStyleBlock(Module):
    """e4e.py:22-46."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c = out_c
        self.spatial = spatial
        num_pools = int(np.log2(spatial))
        modules = [Conv2d(in_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        for _ in range(num_pools - 1):
            modules += [Conv2d(out_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        self.convs = nn.Sequential(*modules)
        self.linear = FullyConnectedLayer(in_features=out_c, out_features=out_c, bias=True, activation='linear', lr_multiplier=1)

    def run_nhwc(self, x):
        """x [B,S,S,in_c] -> [B,out_c]: each stride-2 convolution is evaluated at its input resolution and read back through
        a [::2, ::2] view; its bias + LeakyReLU(0.01) are folded into the operand builder of the next convolution."""
        src, shift, lrelu = x, None, 1.0
        for m in self.convs:
            if not isinstance(m, Conv2d):
                continue
            pack = rt.ConvPack.current(m, '_ia_pack', m.weight, need_wsq=False)
            a, _ = rt.enc_prep([src], shift=shift, lrelu=lrelu, C_pad=pack.Cin_pad)
            src, shift, lrelu = rt.enc_conv(a, m)[:, ::2, ::2], m.bias, 0.01
        _, y = rt.enc_prep([src], shift=shift, lrelu=lrelu, want_split=False, want32=True)
        return self.linear(y.reshape(-1, self.out_c))

    def forward(self, x):
        return self.run_nhwc(_nhwc(x))


class Encoder4Editing(Module):
    """e4e.py:68-134."""

    def __init__(self, n_styles=18, inp_ch=3):
        super().__init__()
        _make_trunk(self, inp_ch)
        self.styles = nn.ModuleList()
        self.style_count = n_styles
        self.coarse_ind = 3
        self.middle_ind = 7
        for i in range(self.style_count):
            spatial = 16 if i < self.coarse_ind else (32 if i < self.middle_ind else 64)
            self.styles.append(GradualStyleBlock(512, 512, spatial))
        self.latlayer1 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)

    def run_nhwc(self, x, w_offset=None):
        """x [B,256,256,inp_ch]-indexed view -> w [B,n_styles,512]; ``w_offset`` [512] (latent_avg) is added to every row."""
        c3, f = _run_trunk(self, x, taps=(6, 20, 23))
        c1, c2 = f[6], f[20]
        B = c3.shape[0]
        w = torch.empty((B, self.style_count, 512), dtype=torch.float32, device=c3.device)

        def row(i):
            return w[:, i].view(B, 1, 1, 512)          # [B,1,1,512] destination with pixel stride n_styles*512
        w0 = self.styles[0].run_nhwc(c3)
        rt.enc_affine_act(w0.view(B, 1, 1, 512), shift=w_offset, out=row(0))
        p2 = rt.enc_upsample_add(c3, _conv_bias_act([c2], self.latlayer1)) if self.style_count > self.coarse_ind else None
        p1 = rt.enc_upsample_add(p2, _conv_bias_act([c1], self.latlayer2)) if self.style_count > self.middle_ind else None

        def style(i):
            features = c3 if i < self.coarse_ind else (p2 if i < self.middle_ind else p1)
            delta = self.styles[i].run_nhwc(features)
            rt.enc_affine_act(delta.view(B, 1, 1, 512), res=w0.view(B, 1, 1, 512), res_shift=w_offset, out=row(i))
        # The map2style heads are independent chains of small, latency-bound launches (a few CTAs each): issue them round-robin
        # on side streams and join before w is consumed (IA_E4E_STREAMS=0: one after the other on the caller's stream).
        n_side = int(os.environ.get('IA_E4E_STREAMS', '4'))
        if n_side <= 1 or self.style_count <= 2:
            for i in range(1, self.style_count):
                style(i)
            return w
        cur = torch.cuda.current_stream(c3.device)
        side = rt.side_streams(c3.device, 2 + n_side)[2:]      # (the first two belong to AR_eval_forward's UNet pair)
        for s_ in side:
            s_.wait_stream(cur)
        for i in range(1, self.style_count):
            with torch.cuda.stream(side[(i - 1) % n_side]):
                style(i)
        for s_ in side:
            cur.wait_stream(s_)
        return w

    def forward(self, x):
        return self.run_nhwc(_nhwc(x))


# ---- unet_encoders.py -------------------------------------------------------------------------------------------
class ConvGRU(Module):
    """unet_encoders.py:8-49."""

    def __init__(self, channels, kernel_size=3, padding=1, out_act_prelu=False):
        super().__init__()
        if out_act_prelu:
            raise NotImplementedError('ConvGRU(out_act_prelu=True) is not instantiated by the inversion encoder')
        self.channels = channels
        self.ih = Sequential(nn.Conv2d(channels * 2, channels * 2, kernel_size, padding=padding), nn.Sigmoid())
        self.hh = Sequential(nn.Conv2d(channels * 2, channels, kernel_size, padding=padding), nn.Tanh())

    def step_nhwc(self, x, h):
        """x, h [B,H,W,C] (h may be None = zero state) -> new h."""
        ih, hh = self.ih[0], self.hh[0]
        if h is None:
            h = torch.zeros(x.shape, dtype=torch.float32, device=x.device)
        a, _ = rt.enc_prep([x, h], C_pad=rt.ConvPack.current(ih, '_ia_pack', ih.weight, need_wsq=False).Cin_pad)
        rh, z = rt.enc_gru_gate0(rt.enc_conv(a, ih), ih.bias, h)
        a, _ = rt.enc_prep([x, rh], C_pad=rt.ConvPack.current(hh, '_ia_pack', hh.weight, need_wsq=False).Cin_pad)
        return rt.enc_gru_gate1(rt.enc_conv(a, hh), hh.bias, h, z)

    def run_nhwc(self, x, h, seq2seq=False):
        """x [B,T,H,W,C] or [B,H,W,C]."""
        if x.ndim == 4:
            h = self.step_nhwc(x, h)
            return h, h
        outs = []
        for t in range(x.shape[1]):
            h = self.step_nhwc(x[:, t], h)
            if seq2seq:
                outs.append(h)
        return (torch.stack(outs, dim=1) if seq2seq else h), h

    def forward(self, x, h, seq2seq=False):
        if x.ndim == 5:
            o, h = self.run_nhwc(x.permute(0, 1, 3, 4, 2), None if h is None else rt.to_nhwc(h), seq2seq)
            return (o.permute(0, 1, 4, 2, 3) if seq2seq else _nchw(o)), _nchw(h)
        o, h = self.run_nhwc(_nhwc(x), None if h is None else rt.to_nhwc(h))
        return _nchw(o), _nchw(h)


class DoubleConv(Module):
    """unet_encoders.py:52-67."""

    def __init__(self, in_channels, out_channels, use_instnorm=False):
        super().__init__()
        if use_instnorm:
            raise NotImplementedError('DoubleConv(use_instnorm=True) is not instantiated by the inversion encoder')
        self.double_conv = nn.Sequential(
            nn.BatchNorm2d(in_channels),
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.PReLU(out_channels),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.PReLU(out_channels),
            nn.PReLU(out_channels))

    def run_nhwc(self, srcs):
        """srcs: list of [B,H,W,Ci] sources (tensor or (tensor, pixel_shuffle)) concatenated along channels."""
        bn, c1, p1, c2, p2, p3 = self.double_conv
        sc, sh = rt.enc_bn_fold(bn, srcs)
        a, _ = rt.enc_prep(srcs, scale=sc, shift=sh, C_pad=rt.ConvPack.current(c1, '_ia_pack', c1.weight, need_wsq=False).Cin_pad)
        pad2 = rt.ConvPack.current(c2, '_ia_pack', c2.weight, need_wsq=False).Cin_pad
        if rt.enc_epilogue_fusion():
            a = rt.enc_conv_act(a, c1, pad2, slope=p1.weight)            # bias + PReLU + operand emission inside c1's epilogue
        else:
            raw = rt.enc_conv(a, c1)
            a, _ = rt.enc_prep([raw], shift=c1.bias, slope=p1.weight, C_pad=pad2)
        return rt.enc_affine_act(rt.enc_conv(a, c2), shift=c2.bias, slope1=p2.weight, slope2=p3.weight)

    def forward(self, x):
        return _nchw(self.run_nhwc([_nhwc(x)]))


class Up(Module):
    """unet_encoders.py:70-82."""

    def __init__(self, in_channels, out_channels, upscale_factor=2):
        super().__init__()
        self.up = nn.PixelShuffle(upscale_factor=upscale_factor)
        self.conv = DoubleConv(in_channels, out_channels)

    def run_nhwc(self, x1, x2):
        return self.conv.run_nhwc([x2, (x1, self.up.upscale_factor)])

    def forward(self, x1, x2):
        return _nchw(self.run_nhwc(_nhwc(x1), _nhwc(x2)))


class recurrent_Up(Module):
    """unet_encoders.py:85-98."""

    def __init__(self, in_channels, out_channels, upscale_factor=2):
        super().__init__()
        self.up = nn.PixelShuffle(upscale_factor=upscale_factor)
        self.conv = DoubleConv(in_channels, out_channels, use_instnorm=False)
        self.conv_gru = ConvGRU(out_channels, out_act_prelu=False)

    def run_nhwc(self, x1, x2, T, r=None, seq2seq=False):
        """x1 [B*T,h,w,C1] (read through PixelShuffle), x2 [B*T,H,W,C2]; r: GRU state [B,H,W,Cout] or None."""
        x = self.conv.run_nhwc([x2, (x1, self.up.upscale_factor)])
        return self.conv_gru.run_nhwc(x.unflatten(0, (-1, T)), r, seq2seq)

    def forward(self, x1, x2, T, r=None, seq2seq=False):
        o, r = self.run_nhwc(_nhwc(x1), _nhwc(x2), T, None if r is None else rt.to_nhwc(r), seq2seq)
        return (o.permute(0, 1, 4, 2, 3) if o.ndim == 5 else _nchw(o)), _nchw(r)


def _repeat_T(t, T):
    """[B,H,W,C] -> [B*T,H,W,C] (the `unsqueeze(1).expand(-1,T,...).flatten(0,1)` of unet_encoders.py:219; a stride-0
    view when B == 1)."""
    B = t.shape[0]
    return t.unsqueeze(1).expand(B, T, *t.shape[1:]).reshape(B * T, *t.shape[1:])


def _state_in(r):
    return None if r is None else rt.to_nhwc(r)


class _UNetBase(Module):
    def _build(self, inp_ch, res, use_gru):
        self.res = res
        self.use_gru = use_gru
        self.face_pool = None if res is None else torch.nn.AdaptiveAvgPool2d((res, res))
        _make_trunk(self, inp_ch)
        up = recurrent_Up if use_gru else Up
        self.up1 = up(1024, 512, upscale_factor=1)
        self.up2 = up(384, 384)
        self.up3 = up(224, 256)
        self.up4 = up(128, 96)

    def _trunk_decoder(self, x, r_list):
        """x: [B,T,C,H,W] or [B,C,H,W] -> ((t1..t4) NHWC decoder outputs, r_list NCHW views, T)."""
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
            return (t1, t2, t3, t4), None, T
        r_list = [None] * 4 if r_list is None else list(r_list)
        t1, r0 = self.up1.run_nhwc(x, c3, T, _state_in(r_list[0]))
        t2, r1 = self.up2.run_nhwc(_repeat_T(t1, T), c2, T, _state_in(r_list[1]))
        t3, r2 = self.up3.run_nhwc(_repeat_T(t2, T), c1, T, _state_in(r_list[2]))
        t4, r3 = self.up4.run_nhwc(_repeat_T(t3, T), c0, T, _state_in(r_list[3]))
        return (t1, t2, t3, t4), [_nchw(r) for r in (r0, r1, r2, r3)], T


class TriPlanefeat_Encoder(_UNetBase):
    """Texture UNet, unet_encoders.py:101-246."""

    def __init__(self, inp_ch, seq2seq=False, res=None, use_gru=False):
        super().__init__()
        self.seq2seq = seq2seq
        self._build(inp_ch, res, use_gru)
        self.outconv0 = nn.Conv2d(384, 32, kernel_size=1, padding=0)
        self.outconv1 = nn.Conv2d(384, 512, kernel_size=1, padding=0)
        self.outconv2 = nn.Conv2d(256, 512, kernel_size=1, padding=0)
        self.outconv3 = nn.Conv2d(96, 256, kernel_size=1, padding=0)

    def forward(self, x, r_list=None, return_list=True):
        (t1, t2, t3, t4), r_list, _ = self._trunk_decoder(x, r_list)
        out_list = [_nchw(_conv_bias_act([t], conv)) for t, conv in ((t2, self.outconv0), (t2, self.outconv1), (t3, self.outconv2),
                                                                    (t4, self.outconv3))]
        return (out_list, r_list) if self.use_gru else out_list


def _sft_head(mod, res, t):
    """stack([condition_scale(t), condition_shift(t)]) -> logical [2,B,C,H,W] over one NHWC buffer (unet_encoders.py:337-345).  The two
    branches share one operand of t; each first convolution applies its bias + LeakyReLU(0.2) in its epilogue and emits the second
    convolution's operand."""
    B, H, W, _ = t.shape
    Cs = getattr(mod, f'condition_scale{res}')[2].out_channels
    out = torch.empty((2, B, H, W, Cs), dtype=torch.float32, device=t.device)
    fuse = rt.enc_epilogue_fusion()
    a = None
    for k, kind in enumerate(('scale', 'shift')):
        c0, _, c2 = getattr(mod, f'condition_{kind}{res}')
        if fuse and c0.out_channels % 4 == 0:
            if a is None:
                a, _ = rt.enc_prep([t], C_pad=rt.ConvPack.current(c0, '_ia_pack', c0.weight, need_wsq=False).Cin_pad)
            a2 = rt.enc_conv_act(a, c0, rt.ConvPack.current(c2, '_ia_pack', c2.weight, need_wsq=False).Cin_pad, lrelu=0.2)
            rt.enc_affine_act(rt.enc_conv(a2, c2), shift=c2.bias, out=out[k])
        else:
            y = _conv_bias_act([t], c0, lrelu=0.2)
            _conv_bias_act([y], c2, out=out[k])
    return out.permute(0, 1, 4, 2, 3)


def _sft_heads(mod, ts):
    """final_head + the five condition_scale / condition_shift heads of a tri-plane SFT UNet (unet_encoders.py:337-345).  The heads
    of the four decoder levels are independent two-convolution chains: they are issued on side streams while the caller's stream
    runs final_head and the 256^2 head, and joined before the dict is returned (IA_SFT_STREAMS=0: all on the caller's stream)."""
    t1, t2, t3, t4 = ts
    dev = t4.device
    out = {}
    side = []
    if os.environ.get('IA_SFT_STREAMS', '1') != '0':
        cur = torch.cuda.current_stream(dev)
        side = rt.side_streams(dev, 11)[7:11]
        for s_ in side:
            s_.wait_stream(cur)
    for k, (res, t) in enumerate(zip((16, 32, 64, 128), ts)):
        if side:
            with torch.cuda.stream(side[k % len(side)]):
                out[res] = mod._head(res, t)
        else:
            out[res] = mod._head(res, t)
    f0, p0, f2, p2 = mod.final_head
    y = _conv_bias_act([(t4, mod.head.upscale_factor)], f0, slope=p0.weight)
    t5 = _conv_bias_act([y], f2, slope=p2.weight)
    out[256] = mod._head(256, t5)
    if side:
        for s_ in side:
            cur.wait_stream(s_)
        for res in (16, 32, 64, 128):
            out[res].record_stream(cur)      # produced on a side stream, consumed (and eventually freed) on the caller's stream
    return {res: out[res] for res in (16, 32, 64, 128, 256)}


class TriPlaneSFTfeat_Encoder(_UNetBase):
    """Tri-plane SFT UNet, unet_encoders.py:249-362."""

    def __init__(self, inp_ch, sft_half=True, res=None, use_gru=False):
        super().__init__()
        self.sft_half = sft_half
        self._build(inp_ch, res, use_gru)
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
        (t1, t2, t3, t4), r_list, _ = self._trunk_decoder(x, r_list)
        out = _sft_heads(self, (t1, t2, t3, t4))
        return (out, r_list) if self.use_gru else out


# ---- uvnet.py -----------------------------------------------------------------------------------------------------
class unet_encoder(nn.Module):
    """uvnet.py:15-23."""

    def __init__(self, encoding_texture=False, encoding_triplane=False):
        super().__init__()
        self.texture_unet = TriPlanefeat_Encoder(inp_ch=7, res=256, use_gru=True) if encoding_texture else None
        self.triplane_unet = TriPlaneSFTfeat_Encoder(inp_ch=6, res=256, use_gru=True) if encoding_triplane else None

    def forward(self, x):
        raise NotImplementedError


class inversionNet(nn.Module):
    """uvnet.py:26-209 (inference surface: encode, get_unet_uvinput, AR_eval_forward, forward)."""

    def __init__(self, G_kwargs=None, generator=None, encoding_texture=True, encoding_triplane=False):
        super().__init__()
        self.face_pool = torch.nn.AdaptiveAvgPool2d((256, 256))
        if generator is not None:
            self.generator = generator
        else:
            from .triplane import TriPlaneGenerator
            kw = {k: v for k, v in dict(G_kwargs).items() if k != 'class_name'}
            self.generator = TriPlaneGenerator(**kw).train().requires_grad_(False)
        self.register_buffer('latent_avg', self.generator.backbone.mapping.w_avg.reshape(1, 512))
        self.n_styles = self.generator.texture_backbone.num_ws
        self.encoder = self.set_encoder(self.n_styles, inp_ch=3)
        self.unet_encoder = unet_encoder(encoding_texture=encoding_texture, encoding_triplane=encoding_triplane)
        self.register_buffer('black_uv_bg', -1 * torch.ones(1, 3, 256, 256, dtype=torch.float32))

    def set_encoder(self, n_styles, inp_ch):
        return Encoder4Editing(n_styles, inp_ch)

    def switch_grad(self, nerf_requires_grad=False):
        for _i in range(self.encoder.middle_ind):
            for p in self.encoder.styles[_i].parameters():
                p.requires_grad = nerf_requires_grad

    def encode(self, x):
        """[B,3,S,S] -> W+ codes [B,n_styles,512] (uvnet.py:107-115)."""
        xn = _face_pool(_nhwc(x.float()), 256)
        return self.encoder.run_nhwc(xn, w_offset=self.latent_avg.reshape(-1))

    def _side_streams(self, device):
        return rt.side_streams(device)

    def _delta(self, y_hat, image):
        """y_hat - image[:, :3] as fp32 NHWC [T,H,W,3] (uvnet.py:181)."""
        neg = self.__dict__.get('_ia_neg1')
        if neg is None or neg.device != y_hat.device:
            neg = torch.full((3,), -1.0, dtype=torch.float32, device=y_hat.device)
            self.__dict__['_ia_neg1'] = neg
        return rt.enc_affine_act(_nhwc(y_hat.float()), res=_nhwc(image[:, :3].float()), res_scale=neg)

    def _uvinput(self, uv, delta_nhwc):
        uv = uv.float()
        T = uv.shape[0]
        pverts = _nhwc(uv[:, 3:6]).contiguous()                       # [T,256,256,3]: (u, v, mask)
        uv_delta = rt.grid_sample_nhwc(delta_nhwc, pverts)            # grid = first two channels
        bg = _nhwc(self.black_uv_bg).contiguous().expand(T, -1, -1, -1)
        uv_delta = rt.lerp_alpha(uv_delta, bg, pverts[..., 2])
        _, x_in = rt.enc_prep([_nhwc(uv[:, 0:3]), uv_delta, pverts[..., 2:3]], want_split=False, want32=True)
        return x_in                                                    # [T,256,256,7]

    def get_unet_uvinput(self, uv, delta_x):
        """uvnet.py:117-121."""
        return _nchw(self._uvinput(uv, rt.to_nhwc(delta_x)))

    def _add_offsets(self, feats, offsets):
        out = [_nchw(rt.enc_affine_act(_nhwc(f.float()), res=_nhwc(o))) for f, o in zip(feats, offsets)]
        return out + list(feats[len(offsets):])

    @torch.no_grad()
    def AR_eval_forward(self, x, vid_c, vid_v, ws, r_list, e4e_results=None, return_fake=False):
        """uvnet.py:160-203."""
        G = self.generator
        T = vid_c.shape[0]
        if ws is None:
            ws = self.encode(x['image'][0:1])
        if e4e_results is None:
            texture_feats = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
            static_feats = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        else:
            texture_feats, static_feats = e4e_results['texture'], e4e_results['static']
        vid_ws = ws.expand(T, -1, -1)
        y_hat_e4e = G.synthesis_withTexture(vid_ws, [f.expand(T, -1, -1, -1) for f in texture_feats], vid_c, vid_v,
                                            static_feats=[f.expand(T, -1, -1, -1) for f in static_feats], noise_mode='const')
        delta = self._delta(y_hat_e4e['image'], x['image'])
        x_input = self._uvinput(x['uv'], delta)
        _, tri_in = rt.enc_prep([_nhwc(x['image'][:, :3].float()), delta], want_split=False, want32=True)

        r_list = [None, None] if r_list is None else list(r_list)
        # The two UNets are independent and made of small, latency-bound launches (few CTAs each): run them side by side on
        # two streams and join before their results are consumed.
        cur = torch.cuda.current_stream(x_input.device)
        s_tex, s_tri = self._side_streams(x_input.device)
        s_tex.wait_stream(cur)
        s_tri.wait_stream(cur)
        with torch.cuda.stream(s_tex):
            texture_offsets, r_list[0] = self.unet_encoder.texture_unet(_nchw(x_input).unsqueeze(0), r_list=r_list[0], return_list=True)
        with torch.cuda.stream(s_tri):
            triplane_feat_offsets, r_list[1] = self.unet_encoder.triplane_unet(_nchw(tri_in).unsqueeze(0), r_list=r_list[1])
        cur.wait_stream(s_tex)
        cur.wait_stream(s_tri)
        for t in list(texture_offsets) + list(r_list[0]) + list(r_list[1]) + list(triplane_feat_offsets.values()):
            t.record_stream(cur)      # produced on a side stream, consumed (and eventually freed) on the caller's stream
        texture_feats = self._add_offsets(texture_feats, texture_offsets)
        static_feats = G.backbone.synthesis(ws, cond_list=None, return_list=True, feat_conditions=triplane_feat_offsets,
                                            update_emas=False, noise_mode='const')
        updated = {'w': ws, 'texture': texture_feats, 'static': static_feats}
        if not return_fake:
            return updated, r_list
        fake = G.synthesis_withTexture(vid_ws, [f.expand(T, -1, -1, -1) for f in updated['texture']], vid_c, vid_v,
                                       static_feats=[f.expand(T, -1, -1, -1) for f in updated['static']], noise_mode='const',
                                       evaluation=True)['image']
        return updated, {'e4e': y_hat_e4e['image'], 'image': fake, 'x_input': _nchw(x_input)}, r_list

    def forward(self, x, cam, v, e4e_results=None, visualize_input=False, return_feats=False):
        """uvnet.py:123-157 (single-frame path: the UNets run with T = 1 per sample)."""
        G = self.generator
        if e4e_results is None:
            ws = self.encode(x['image'][:, :3])
            tex = G.texture_backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
            sta = G.backbone.synthesis(ws, cond_list=None, return_list=True, update_emas=False, noise_mode='const')
        else:
            ws, tex, sta = e4e_results['w'], e4e_results['texture'], e4e_results['static']
        y_hat = G.synthesis_withTexture(ws, tex, cam, v, static_feats=sta, noise_mode='const')
        if y_hat['image'].shape[-1] != x['image'].shape[-1]:
            raise NotImplementedError('inversionNet.forward: generator and input image resolutions must agree')
        delta = self._delta(y_hat['image'], x['image'])
        x_input = self._uvinput(x['uv'], delta)
        offsets = self.unet_encoder.texture_unet(_nchw(x_input), return_list=True)
        offsets = offsets[0] if isinstance(offsets, tuple) else offsets
        texture_feats = self._add_offsets(tex, offsets)
        _, tri_in = rt.enc_prep([_nhwc(x['image'][:, :3].float()), delta], want_split=False, want32=True)
        sft = self.unet_encoder.triplane_unet(_nchw(tri_in))
        sft = sft[0] if isinstance(sft, tuple) else sft
        static_feats = G.backbone.synthesis(ws, cond_list=None, return_list=True, feat_conditions=sft, update_emas=False, noise_mode='const')
        output = G.synthesis_withTexture(ws, texture_feats, cam, v, static_feats=static_feats, noise_mode='const')
        if return_feats:
            output['texture'], output['static'] = texture_feats, static_feats
        output['w'] = ws
        output['e4e_image'] = y_hat['image']
        if visualize_input:
            output['x_input'] = _nchw(x_input).clamp(-1, 1)
        return output
