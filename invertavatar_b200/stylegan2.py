"""StyleGAN2 generator blocks of the Next3D++ generator on the B200 engine.

Same classes, constructor arguments, parameter/buffer names and forward signatures as the reference
(training_avatar_texture/networks_stylegan2_new.py: FullyConnectedLayer :96, MappingNetwork :193, SynthesisLayer :276,
ToRGBLayer :340, SynthesisBlock :365, SynthesisNetwork :475, Generator :559; training/networks_stylegan2.py is the same
arithmetic without the cond/feat hooks), so ``copy_params_and_buffers(require_all=True)`` works in both directions and
parameters drawn under ``torch.manual_seed`` match the reference's.

Underneath, nothing is shared with the reference: a layer is (styles + demod coefficients) -> (modulate + bf16 hi/lo
split) -> tcgen05 implicit-GEMM convolution with a fused demod/noise/bias/lrelu/clamp epilogue -- the reference's own
``fused_modconv=False`` formulation (networks_stylegan2_new.py:70-79), so the batch folds into GEMM-M.  Activations
stay channels-last between layers; tensors handed to the caller are logical NCHW views of those buffers.
Forward only (the inference scripts run under no_grad / requires_grad_(False))."""
import math

import numpy as np
import torch

from . import persistence
from . import runtime as rt

SQRT2 = math.sqrt(2.0)


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR filter setup, reference torch_utils/ops/upfirdn2d.py:72-116."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _is_1331(f):
    """True when the resample filter is the stock outer([1,3,3,1])/64 the fused kernels assume."""
    if f is None or f.ndim != 2 or tuple(f.shape) != (4, 4):
        return False
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    ref = torch.outer(k, k) / 64.0
    return bool(torch.allclose(f.detach().float().cpu(), ref, atol=1e-7))


@persistence.persistent_class
class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        return rt.fully_connected(x, self.weight, self.bias, w_gain=self.weight_gain, b_gain=self.bias_gain,
                                  act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


@persistence.persistent_class
class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(features[idx], features[idx + 1], activation=activation,
                                                          lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        x = None
        if self.z_dim > 0:
            assert z.shape[1] == self.z_dim
            x = rt.normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            assert c.shape[1] == self.c_dim
            y = rt.normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas and self.w_avg_beta is not None:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is None:
            if truncation_psi != 1:
                x = self.w_avg.lerp(x, truncation_psi)
            return x
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
        return rt.broadcast_truncate(x, getattr(self, 'w_avg', None), self.num_ws, psi=truncation_psi,
                                     cutoff=truncation_cutoff)

    def extra_repr(self):
        return f'z_dim={self.z_dim:d}, c_dim={self.c_dim:d}, w_dim={self.w_dim:d}, num_ws={self.num_ws:d}'


def _noise_for(layer, noise_mode, batch, device):
    """Per-layer noise image(s): 'const' -> the registered buffer [R,R]; 'random' -> fresh N(0,1) [B,R,R]
    (networks_stylegan2_new.py:317-321).  The multiplication by noise_strength happens inside the epilogue."""
    if not layer.use_noise or noise_mode == 'none':
        return None
    if noise_mode == 'const':
        return layer.noise_const
    return torch.randn([batch, layer.resolution, layer.resolution], device=device)


@persistence.persistent_class
class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim = in_channels, out_channels, w_dim
        self.resolution, self.up, self.use_noise = resolution, up, use_noise
        self.activation, self.conv_clamp = activation, conv_clamp
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = rt.ACT_DEFAULTS[activation][1]
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    # ---- engine-facing pieces ----
    def fmt(self):
        """Operand format this layer's convolution runs in (rt.layer_fmt: the owner's per-layer precision choice)."""
        return rt.layer_fmt(self)

    def pack(self):
        return rt.ConvPack.current(self, '_ia_pack', self.weight, need_wsq=True, fmt=self.fmt())

    def style_entry(self, w_index):
        p = self.pack()
        return dict(affine_w=self.affine.weight, affine_b=self.affine.bias, wsq=p.wsq, Cin=self.in_channels,
                    Cout=self.out_channels, w_index=w_index, style_gain=1.0)

    def run_nhwc(self, x, styles, dcoef, noise_mode='const', gain=1, cond=None, cond_alpha=None):
        """x [B,H,W,Cin] fp32 NHWC -> [B,R,R,Cout] fp32 NHWC."""
        assert self.up in (1, 2), 'only up=1 and up=2 synthesis layers exist on the hot path'
        if self.up == 2 and not _layer_filter_ok(self):
            raise RuntimeError('SynthesisLayer: the fused up=2 path assumes the [1,3,3,1] resample filter')
        B, H, W, Cin = x.shape
        assert Cin == self.in_channels and H * self.up == self.resolution, (x.shape, self.in_channels, self.resolution)
        pack = self.pack()
        hi, lo = rt.modsplit(x, styles, cond=cond, cond_alpha=cond_alpha, C_pad=pack.Cin_pad, fmt=pack.fmt)
        noise = _noise_for(self, noise_mode, B, x.device)
        strength = self.noise_strength if noise is not None else None
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        R = self.resolution
        out = torch.empty((B, R, R, self.out_channels), dtype=torch.float32, device=x.device)
        if self.up == 1:
            rt.conv_same(hi, lo, pack, pack.Cin_pad, out, dcoef=dcoef, noise=noise, noise_strength=strength, bias=self.bias,
                         act=self.activation, gain=act_gain, clamp=act_clamp, mode=1)
        else:
            raw = torch.empty((B, 2 * H + 1, 2 * W + 1, self.out_channels), dtype=torch.float32, device=x.device)
            rt.conv_transpose_up2_raw(hi, lo, pack, pack.Cin_pad, raw)
            rt.fir_epilogue(raw, rt.fir4x4_gain4(x.device), out, dcoef, noise, strength, self.bias, self.activation,
                            act_gain, act_clamp)
        return out

    def run_split(self, a, dcoef, noise_mode='const', gain=1, want32=False, e1=None, e2=None, rgb=None):
        """Fused-chain entry: ``a`` is an rt.Split already holding x*styles of THIS layer; the epilogue writes any of
        the fp32 NHWC result (want32) and the pre-modulated operands of the consumers (e1/e2 = (Split, styles))."""
        assert self.up in (1, 2)
        if self.up == 2 and not _layer_filter_ok(self):
            raise RuntimeError('SynthesisLayer: the fused up=2 path assumes the [1,3,3,1] resample filter')
        B, H, W, _ = a.hi.shape
        pack = self.pack()
        assert a.C_pad == pack.Cin_pad and H * self.up == self.resolution
        noise = _noise_for(self, noise_mode, B, a.hi.device)
        strength = self.noise_strength if noise is not None else None
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        R = self.resolution
        out = torch.empty((B, R, R, self.out_channels), dtype=torch.float32, device=a.hi.device) if want32 else None
        if self.up == 1:
            assert a.img_pix == 0, 'a row-padded operand feeds transposed convolutions only'
            rt.conv_same(a.hi, a.lo, pack, pack.Cin_pad, out, dcoef=dcoef, noise=noise, noise_strength=strength, bias=self.bias,
                         act=self.activation, gain=act_gain, clamp=act_clamp, mode=1, e1=e1, e2=e2, rgb=rgb)
        else:
            assert rgb is None
            raw = torch.empty((B, 2 * H + 1, 2 * W + 1, self.out_channels), dtype=torch.float32, device=a.hi.device)
            rt.conv_transpose_up2_raw(a.hi, a.lo, pack, pack.Cin_pad, raw, img_rows=a.img_rows)
            rt.fir_epilogue(raw, rt.fir4x4_gain4(a.hi.device), out, dcoef, noise, strength, self.bias, self.activation,
                            act_gain, act_clamp, e1=e1, e2=e2)
        return out

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
        assert noise_mode in ['random', 'const', 'none']
        plan = _plan_for(self, [self.style_entry(0)])
        styles, dcoefs = plan.run(w.reshape(w.shape[0], 1, -1))
        y = self.run_nhwc(rt.to_nhwc(x), styles[0], dcoefs[0], noise_mode=noise_mode, gain=gain)
        return rt.from_nhwc(y)

    def extra_repr(self):
        return ' '.join([f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d},',
                         f'resolution={self.resolution:d}, up={self.up}, activation={self.activation:s}'])


def _layer_filter_ok(layer):
    ok = layer.__dict__.get('_ia_filter_ok')
    if ok is None:
        ok = _is_1331(layer.resample_filter)
        layer.__dict__['_ia_filter_ok'] = ok
    return ok


def _plan_for(owner, entries, attr='_ia_plan', share=None):
    """StylePlan cached on ``owner``; rebuilt when any referenced tensor moved (new storage / device)."""
    key = tuple((e['affine_w'].data_ptr(), e['affine_b'].data_ptr(), 0 if e['wsq'] is None else e['wsq'].data_ptr(),
                 e['w_index']) for e in entries) + (None if share is None else tuple(share),)
    cached = owner.__dict__.get(attr)
    if cached is None or cached[1] is None or cached[0] != key:
        cached = (key, rt.StylePlan(entries, entries[0]['affine_w'].device, share=share))
        owner.__dict__[attr] = cached
    return cached[1]


@persistence.persistent_class
class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim = in_channels, out_channels, w_dim
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def pack(self):
        return rt.ConvPack.current(self, '_ia_pack', self.weight, need_wsq=False)

    def style_entry(self, w_index):
        return dict(affine_w=self.affine.weight, affine_b=self.affine.bias, wsq=None, Cin=self.in_channels,
                    Cout=self.out_channels, w_index=w_index, style_gain=float(self.weight_gain))

    def weight2d(self):
        """[out_channels, in_channels] fp32 view of the 1x1 kernel (operand of the fused ToRGB contraction)."""
        return self.weight.detach().reshape(self.out_channels, self.in_channels)

    def run_nhwc(self, x, styles, img_prev=None, out_nchw=False):
        """x [B,H,W,Cin] -> img = upsample2d(img_prev) + clamp(modconv1x1(x) + bias)  as NHWC (or planar NCHW)."""
        B, H, W, _ = x.shape
        pack = self.pack()
        hi, lo = rt.modsplit(x, styles, C_pad=pack.Cin_pad)
        raw = torch.empty((B, H, W, self.out_channels), dtype=torch.float32, device=x.device)
        rt.conv_same(hi, lo, pack, pack.Cin_pad, raw, mode=0)
        return rt.torgb_finish(raw, self.bias, self.conv_clamp, img_prev, out_nchw=out_nchw)

    def run_split(self, a, img_prev=None, out_nchw=False):
        """``a``: rt.Split holding x*styles of this layer (emitted by the producing convolution's epilogue)."""
        B, H, W, _ = a.hi.shape
        pack = self.pack()
        if not out_nchw and rt.can_fuse_torgb_tail(H, W, self.out_channels):
            # bias, clamp and the upsampled previous image are applied by the 1x1 convolution's epilogue (mode 2): no raw
            # round trip, no ToRGB-tail launch; same arithmetic as ia_torgb_finish
            img = torch.empty((B, H, W, self.out_channels), dtype=torch.float32, device=a.hi.device)
            rt.conv_same(a.hi, a.lo, pack, pack.Cin_pad, img, bias=self.bias, clamp=self.conv_clamp, mode=2, img_prev=img_prev)
            return img
        raw = torch.empty((B, H, W, self.out_channels), dtype=torch.float32, device=a.hi.device)
        rt.conv_same(a.hi, a.lo, pack, pack.Cin_pad, raw, mode=0)
        return rt.torgb_finish(raw, self.bias, self.conv_clamp, img_prev, out_nchw=out_nchw)

    def forward(self, x, w, fused_modconv=True):
        plan = _plan_for(self, [self.style_entry(0)])
        styles, _ = plan.run(w.reshape(w.shape[0], 1, -1))
        return rt.from_nhwc(self.run_nhwc(rt.to_nhwc(x), styles[0]))

    def extra_repr(self):
        return f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d}'


@persistence.persistent_class
class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=256, use_fp16=False, fp16_channels_last=False,
                 fused_modconv_default=True, **layer_kwargs):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        if architecture == 'resnet':
            raise NotImplementedError('resnet synthesis blocks are not on the generator hot path (skip architecture only)')
        self.in_channels, self.w_dim, self.resolution = in_channels, w_dim, resolution
        self.img_channels, self.is_last, self.architecture = img_channels, is_last, architecture
        self.use_fp16 = use_fp16
        self.channels_last = (use_fp16 and fp16_channels_last)
        self.fused_modconv_default = fused_modconv_default
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp,
                                        channels_last=self.channels_last, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == 'skip':
            self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last)
            self.num_torgb += 1

    def _plan(self):
        entries = []
        if self.in_channels != 0:
            entries.append(self.conv0.style_entry(len(entries)))
        entries.append(self.conv1.style_entry(len(entries)))
        if hasattr(self, 'torgb'):
            entries.append(self.torgb.style_entry(len(entries)))
        return _plan_for(self, entries)

    def run_nhwc(self, x, img, ws, condition=None, noise_mode='random', cond=None, cond_alpha=None, img_nchw=False,
                 gain=1):
        """Channels-last block: x [B,H/2,W/2,Cin] or None, img [B,H/2,W/2,Cimg] or None, ws [B,n,w_dim].
        ``cond``/``cond_alpha`` blend the incoming x (face-backbone cond_list) inside the operand preparation."""
        if self.is_last is False and self.architecture == 'orig':
            pass
        styles, dcoefs = self._plan().run(ws)
        i = 0
        if self.in_channels == 0:
            B = ws.shape[0]
            x = self.const.detach().permute(1, 2, 0).unsqueeze(0).expand(B, -1, -1, -1).contiguous()
            x = self.conv1.run_nhwc(x, styles[i], dcoefs[i], noise_mode=noise_mode, gain=gain)
            i += 1
        else:
            x = self.conv0.run_nhwc(x, styles[i], dcoefs[i], noise_mode=noise_mode, gain=gain, cond=cond, cond_alpha=cond_alpha)
            i += 1
            if condition is not None:
                # CS-SFT (networks_stylegan2_new.py:448-452): second half of the channels <- x*scale + shift.
                # In place on the fp32 NHWC activation (ia_sft_half).
                rt.sft_half(x, condition[0].permute(0, 2, 3, 1), condition[1].permute(0, 2, 3, 1))
            x = self.conv1.run_nhwc(x, styles[i], dcoefs[i], noise_mode=noise_mode, gain=gain)
            i += 1
        if hasattr(self, 'torgb'):
            img = self.torgb.run_nhwc(x, styles[i], img_prev=img, out_nchw=img_nchw)
        elif img is not None:
            img = rt.to_nhwc(rt.upfirdn2d(rt.from_nhwc(img), self.resample_filter, up=(2, 2), padding=(2, 1, 2, 1), gain=4.0))
        return x, img

    def layers(self):
        """[(kind, layer)] in evaluation order."""
        out = []
        if self.in_channels != 0:
            out.append(('conv', self.conv0))
        out.append(('conv', self.conv1))
        if hasattr(self, 'torgb'):
            out.append(('torgb', self.torgb))
        return out

    def run_chain(self, a_in, img, styles, dcoefs, B, noise_mode='const', condition=None, want_x32=False, next_conv=None,
                  next_styles=None, img_nchw=False, gain=1, skip_rgb=False):
        """Fused block: every convolution epilogue writes the operand(s) of its consumer(s) directly.
        a_in: rt.Split for conv0 (None for the const block); styles/dcoefs: this block's entries in layers() order;
        next_conv/next_styles: the following block's conv0 and its styles (None -> no operand emitted for it).
        skip_rgb: this block's image (ToRGB + skip-connection up-sampling) has no consumer -- not computed, img comes back None.
        Returns (x32 or None, img, a_next or None)."""
        dev = styles[0].device
        R = self.resolution
        i = 0
        if self.in_channels == 0:
            x0 = self.const.detach().permute(1, 2, 0).unsqueeze(0).expand(B, -1, -1, -1).contiguous()
            hi, lo = rt.modsplit(x0, styles[0], C_pad=self.conv1.pack().Cin_pad, fmt=self.conv1.fmt())
            a1 = rt.Split(hi, lo)
        else:
            a1 = rt.new_split(B, R, R, self.conv1.pack().Cin_pad, dev, C=self.conv1.in_channels, fmt=self.conv1.fmt())
            if condition is None:
                self.conv0.run_split(a_in, dcoefs[0], noise_mode=noise_mode, gain=gain, e1=(a1, styles[1]))
            else:
                # CS-SFT (networks_stylegan2_new.py:448-452) needs the fp32 activation: once per identity, not per frame
                x = self.conv0.run_split(a_in, dcoefs[0], noise_mode=noise_mode, gain=gain, want32=True)
                rt.sft_half(x, condition[0].permute(0, 2, 3, 1), condition[1].permute(0, 2, 3, 1))
                hi, lo = rt.modsplit(x, styles[1], C_pad=self.conv1.pack().Cin_pad, fmt=self.conv1.fmt())
                a1 = rt.Split(hi, lo)
            i = 1
        has_rgb = hasattr(self, 'torgb') and not skip_rgb
        if skip_rgb:
            img = None
        # ToRGB with few image channels (the super-resolution blocks: 3) is contracted inside conv1's epilogue: no ToRGB operand is
        # written and re-read, no 1x1 convolution launch
        fuse_rgb = (has_rgb and not want_x32 and self.torgb.weight.shape[2] == 1 and
                    rt.can_fuse_torgb(R, R, self.conv1.out_channels, self.torgb.out_channels))
        a_rgb = rt.new_split(B, R, R, self.torgb.pack().Cin_pad, dev, C=self.torgb.in_channels) if (has_rgb and not fuse_rgb) else None
        a_next = rt.new_split(B, R, R, next_conv.pack().Cin_pad, dev, C=next_conv.in_channels,
                              pad_row=next_conv.up == 2 and rt.pad_row_wanted(R, R), fmt=next_conv.fmt()) if next_conv is not None else None
        rgb_raw = torch.zeros((B, R, R, self.torgb.out_channels), dtype=torch.float32, device=dev) if fuse_rgb else None
        x32 = self.conv1.run_split(a1, dcoefs[i], noise_mode=noise_mode, gain=gain, want32=want_x32,
                                   e1=(a_next, next_styles) if a_next is not None else None,
                                   e2=(a_rgb, styles[i + 1]) if a_rgb is not None else None,
                                   rgb=(rgb_raw, self.torgb.weight2d(), styles[i + 1]) if fuse_rgb else None)
        if fuse_rgb:
            img = rt.torgb_finish(rgb_raw, self.torgb.bias, self.torgb.conv_clamp, img, out_nchw=img_nchw)
        elif has_rgb:
            img = self.torgb.run_split(a_rgb, img_prev=img, out_nchw=img_nchw)
        elif img is not None:
            img = rt.to_nhwc(rt.upfirdn2d(rt.from_nhwc(img), self.resample_filter, up=(2, 2), padding=(2, 1, 2, 1), gain=4.0))
        return x32, img, a_next

    def forward(self, x, img, ws, condition=None, force_fp32=False, fused_modconv=None, update_emas=False, **layer_kwargs):
        assert ws.shape[1] == self.num_conv + self.num_torgb and ws.shape[2] == self.w_dim
        noise_mode = layer_kwargs.get('noise_mode', 'random')
        if self.in_channels != 0:
            assert tuple(x.shape[1:]) == (self.in_channels, self.resolution // 2, self.resolution // 2)
            x = rt.to_nhwc(x)
        if img is not None:
            assert tuple(img.shape[1:]) == (self.img_channels, self.resolution // 2, self.resolution // 2)
            img = rt.to_nhwc(img)
        x, img = self.run_nhwc(x, img, ws, condition=condition, noise_mode=noise_mode, gain=layer_kwargs.get('gain', 1))
        return rt.from_nhwc(x), (rt.from_nhwc(img) if img is not None else None)

    def extra_repr(self):
        return f'resolution={self.resolution:d}, architecture={self.architecture:s}'


@persistence.persistent_class
class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4,
                 **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.num_fp16_res = num_fp16_res
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res // 2] if res > 4 else 0
            out_channels = channels_dict[res]
            use_fp16 = (res >= fp16_resolution)
            is_last = (res == self.img_resolution)
            block = SynthesisBlock(in_channels, out_channels, w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=is_last, use_fp16=use_fp16, **block_kwargs)
            self.num_ws += block.num_conv
            if is_last:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def style_pass(self, ws):
        """One style/demod pass for the whole network (2 launches instead of 2 per block) -> (styles, dcoefs, spans):
        per-layer lists in evaluation order and the [first, last) span of every block."""
        blocks = [getattr(self, f'b{res}') for res in self.block_resolutions]
        entries, spans, w_idx = [], [], 0
        for block in blocks:
            lay = block.layers()
            first = len(entries)
            for j, (kind, layer) in enumerate(lay):
                entries.append(layer.style_entry(w_idx + j))
            spans.append((first, len(entries)))
            w_idx += block.num_conv
        styles, dcoefs = _plan_for(self, entries).run(ws)
        return styles, dcoefs, spans

    def forward(self, ws, cond_list=None, return_list=False, feat_conditions=None, return_imgs=False, out_res=(32, 256),
                **block_kwargs):
        """cond_list / return_list / feat_conditions semantics of networks_stylegan2_new.py:509-548.  (The stock
        training/networks_stylegan2.SynthesisNetwork.forward(ws, **kw) is the cond_list=None, return_list=False case.)
        Engine-internal kwarg ``prune_after=res`` (return_list only): the caller reads nothing above the features of block
        ``res`` -- the list then ends with that block's x ([img_start, x_start, ..., x_res]); the blocks above it and the
        skip-connection images after the first one are dead code and are not evaluated (the surviving entries are bit-identical
        to the unpruned call's)."""
        assert not (return_list and return_imgs)
        prune_after = block_kwargs.pop('prune_after', None)
        assert prune_after is None or (return_list and cond_list is None), "prune_after applies to plain return_list calls"
        assert ws.shape[1] == self.num_ws and ws.shape[2] == self.w_dim, (ws.shape, self.num_ws)
        noise_mode = block_kwargs.get('noise_mode', 'random')
        ws = ws.to(torch.float32)
        B = ws.shape[0]
        blocks = [getattr(self, f'b{res}') for res in self.block_resolutions]
        prefix = block_kwargs.pop('prefix', None)     # engine-internal: result of synthesis_prefix_grouped for this network
        if prefix is not None:
            styles, dcoefs, spans = prefix['styles'], prefix['dcoefs'], prefix['spans']
        else:
            styles, dcoefs, spans = self.style_pass(ws)
        x_list, out_imgs = [], []
        start_layer = int(np.log2(out_res[0])) - 2
        end_layer = (self.img_resolution_log2 - 2) if len(out_res) == 1 else (int(np.log2(out_res[1])) - 2)
        a = img = None
        for index, (res, block) in enumerate(zip(self.block_resolutions, blocks)):
            lo_i, hi_i = spans[index]
            cond_feat = feat_conditions[res] if (feat_conditions is not None and res in feat_conditions.keys()) else None
            emitting = index >= start_layer
            # the blend x <- cond*alpha + x*(1-alpha) (:538-540) happens between this block and the next one: then the next
            # operand cannot come straight out of this block's epilogue
            blend_next = cond_list is not None and emitting and index < end_layer
            want_x32 = (emitting and (return_list or return_imgs)) or blend_next
            has_next = index + 1 < len(blocks)
            next_conv = blocks[index + 1].conv0 if (has_next and not blend_next) else None
            next_styles = styles[spans[index + 1][0]] if next_conv is not None else None
            if prune_after is not None and res > prune_after:
                break                                     # nothing above this block is read by the caller
            dead_img = prune_after is not None and index > start_layer     # images after the first exported one feed only later images
            if prune_after is not None and res == prune_after:
                next_conv = next_styles = None
            if prefix is not None and index < prefix['index']:
                continue                                  # evaluated by the grouped prefix
            if prefix is not None and index == prefix['index']:
                assert cond_feat is None and index <= start_layer
                x32, img, a = prefix['x32'], prefix['img'], prefix['a_next']
            else:
                x32, img, a = block.run_chain(a, img, styles[lo_i:hi_i], dcoefs[lo_i:hi_i], B, noise_mode=noise_mode, condition=cond_feat,
                                              want_x32=want_x32, next_conv=next_conv, next_styles=next_styles, skip_rgb=dead_img)
            if emitting:
                if return_list:
                    if index == start_layer:
                        x_list.append(rt.from_nhwc(img))
                    x_list.append(rt.from_nhwc(x32))
                if return_imgs:
                    if index == start_layer:
                        out_imgs.append(rt.from_nhwc(x32))
                    out_imgs.append(rt.from_nhwc(img))
                if cond_list is not None:
                    if index == start_layer:
                        c_img, c_a = _split_cond(cond_list[0])
                        img = rt.lerp_alpha(c_img, img, c_a)
                    if blend_next and has_next:
                        cnd, cal = _split_cond(cond_list[1 + index - start_layer])
                        nxt = blocks[index + 1].conv0
                        a = rt.modsplit_split(x32, styles[spans[index + 1][0]], cond=cnd, cond_alpha=cal, C_pad=nxt.pack().Cin_pad,
                                              pad_row=nxt.up == 2 and rt.pad_row_wanted(x32.shape[1], x32.shape[2]), fmt=nxt.fmt())
        if return_list:
            if prune_after is None:
                x_list.append(rt.from_nhwc(img))
            return x_list
        if return_imgs:
            return out_imgs
        return rt.from_nhwc(img)

    def extra_repr(self):
        return ' '.join([f'w_dim={self.w_dim:d}, num_ws={self.num_ws:d},',
                         f'img_resolution={self.img_resolution:d}, img_channels={self.img_channels:d},',
                         f'num_fp16_res={self.num_fp16_res:d}'])


def _split_cond(c):
    """A cond_list entry is either the engine's (features NHWC [B,H,W,C], alpha [B,H,W]) pair or the reference's
    [B,C+1,H,W] tensor whose last channel is the blend alpha (triplane_v20.py:335-337)."""
    if isinstance(c, (tuple, list)):
        return c[0], c[1]
    cn = rt.to_nhwc(c)
    return cn[..., :-1], cn[..., -1].contiguous()


def _stack_cached(owner, attr, tensors, flatten_from=None, pad_to=None):
    """torch.stack of small per-network tensors (bias, noise image, noise strength), cached on ``owner`` until one changes;
    1-D members are zero-padded to ``pad_to`` entries first."""
    key = tuple((t.data_ptr(), t._version) for t in tensors) + (pad_to,)
    hit = owner.__dict__.get(attr)
    if hit is None or hit[0] != key:
        members = [t.detach().float() for t in tensors]
        if pad_to is not None:
            members = [torch.nn.functional.pad(t, (0, pad_to - t.shape[0])) for t in members]
        st = torch.stack(members)
        if flatten_from is not None:
            st = st.reshape(*st.shape[:flatten_from], -1)
        hit = (key, st.contiguous())
        owner.__dict__[attr] = hit
    return hit[1]


def can_group_prefix(nets, upto_res=32):
    """True when the networks have identical layer shapes up to ``upto_res`` (image channels may differ) and the fused
    up-sampling path applies, i.e. synthesis_prefix_grouped may evaluate their low-resolution blocks as one batch."""
    n0 = nets[0]
    for n in nets:
        if n.block_resolutions != n0.block_resolutions or n.w_dim != n0.w_dim:
            return False
        for res in n.block_resolutions:
            if res > upto_res:
                break
            b, b0 = getattr(n, f'b{res}'), getattr(n0, f'b{res}')
            if b.architecture != 'skip' or not hasattr(b, 'torgb') or b.in_channels != b0.in_channels:
                return False
            for (k, l), (k0, l0) in zip(b.layers(), b0.layers()):
                if k != k0 or l.in_channels != l0.in_channels or (k == 'conv' and (l.out_channels != l0.out_channels or l.up != l0.up or
                                                                                 l.activation != l0.activation or l.conv_clamp != l0.conv_clamp or
                                                                                 l.use_noise != l0.use_noise or l.fmt() != l0.fmt() or (l.up == 2 and not _layer_filter_ok(l)))):
                    return False
                if k == 'torgb' and l.conv_clamp != l0.conv_clamp:
                    return False
    return upto_res in n0.block_resolutions and n0.block_resolutions[-1] > upto_res


def synthesis_prefix_grouped(nets, ws, noise_mode='const', upto_res=32):
    """Blocks b4..b{upto_res} of several SynthesisNetworks evaluated as ONE grouped batch (ia_conv_params.groups): the
    low-resolution layers are latency-bound (a few CTAs, a long serial k-loop), so running the three backbones' copies side by
    side costs about the same time as one of them.  Arithmetic per network is unchanged (same kernels, same operand order).
    Returns one ``prefix`` dict per network for ``SynthesisNetwork.forward(..., prefix=...)``."""
    G = len(nets)
    ws = ws.to(torch.float32)
    B = ws.shape[0]
    dev = ws.device
    k_last = nets[0].block_resolutions.index(upto_res)
    # ONE style / demodulation pass for all G networks (2 launches): the layers the grouped launches read -- every layer of the
    # blocks up to upto_res plus the first layer of the block that follows -- write rows [g*B, (g+1)*B) of one group-major buffer
    # per layer, so no concatenation is needed; the remaining layers get their own buffers, handed to each network's forward.
    per_net = []
    for n in nets:
        entries, spans, w_idx = [], [], 0
        for res in n.block_resolutions:
            block = getattr(n, f'b{res}')
            first = len(entries)
            for j, (kind, layer) in enumerate(block.layers()):
                entries.append(layer.style_entry(w_idx + j))
            spans.append((first, len(entries)))
            w_idx += block.num_conv
        per_net.append((entries, spans))
    spans = per_net[0][1]
    n_shared = spans[k_last + 1][0] + 1
    all_entries, share, offs = [], [], []
    for g, (entries, _) in enumerate(per_net):
        offs.append(len(all_entries))
        for i, e in enumerate(entries):
            all_entries.append(e)
            share.append((i, g, G) if i < n_shared else None)
    plan = _plan_for(nets[0], all_entries, attr='_ia_gplan', share=share)
    st_all, dc_all = plan.run(ws)
    shared = plan.shared_for(B)
    passes = [(st_all[o:o + len(per_net[g][0])], dc_all[o:o + len(per_net[g][0])], per_net[g][1]) for g, o in enumerate(offs)]
    group = lambda gstride: (G, B, gstride)

    def cat_layer(idx, which):       # per-layer styles / dcoefs of all networks, group-major (rows g*B .. of the shared buffer)
        return shared[idx][which]

    def layer_noise(layers):
        l0 = layers[0]
        if not l0.use_noise or noise_mode == 'none':
            return None, None, 0
        strength = _stack_cached(l0, '_ia_g_strength', [l.noise_strength for l in layers])
        R = l0.resolution
        if noise_mode == 'const':
            return _stack_cached(l0, '_ia_g_noise', [l.noise_const for l in layers], flatten_from=1), strength, R * R
        return torch.randn([G * B, R, R], device=dev), strength, 0

    a = img = x32 = None
    a_next = None
    for index in range(k_last + 1):
        res = nets[0].block_resolutions[index]
        blocks = [getattr(n, f'b{res}') for n in nets]
        b0 = blocks[0]
        lo_i = spans[index][0]
        i = 0
        conv1s = [b.conv1 for b in blocks]
        pack1 = rt.ConvPackGroup.current(conv1s[0], '_ia_gpack', [c.weight for c in conv1s], fmt=conv1s[0].fmt())
        if b0.in_channels == 0:
            x0 = torch.cat([b.const.detach().permute(1, 2, 0).unsqueeze(0).expand(B, -1, -1, -1) for b in blocks], dim=0).contiguous()
            hi, lo = rt.modsplit(x0, cat_layer(lo_i, 0), C_pad=pack1.Cin_pad, fmt=pack1.fmt)
            a1 = rt.Split(hi, lo)
        else:
            conv0s = [b.conv0 for b in blocks]
            pack0 = rt.ConvPackGroup.current(conv0s[0], '_ia_gpack', [c.weight for c in conv0s], fmt=conv0s[0].fmt())
            a1 = rt.new_split(G * B, res, res, pack1.Cin_pad, dev, C=conv1s[0].in_channels, fmt=pack1.fmt)
            H = res // 2
            raw = torch.empty((G * B, 2 * H + 1, 2 * H + 1, pack0.Cout), dtype=torch.float32, device=dev)
            rt.conv_transpose_up2_raw(a.hi, a.lo, pack0, pack0.Cin_pad, raw, group=group(0))
            nz, ns, gs = layer_noise(conv0s)
            rt.fir_epilogue(raw, rt.fir4x4_gain4(dev), None, cat_layer(lo_i, 1), nz, ns, _stack_cached(conv0s[0], '_ia_g_bias', [c.bias for c in conv0s]),
                            conv0s[0].activation, conv0s[0].act_gain, conv0s[0].conv_clamp, e1=(a1, cat_layer(lo_i + 1, 0)), group=group(gs))
            i = 1
        last = index == k_last
        rgbs = [b.torgb for b in blocks]
        packr = rt.ConvPackGroup.current(rgbs[0], '_ia_gpack', [t.weight for t in rgbs], need_wsq=False)
        a_rgb = rt.new_split(G * B, res, res, packr.Cin_pad, dev, C=rgbs[0].in_channels)
        nxt = [getattr(n, f'b{nets[0].block_resolutions[index + 1]}').conv0 for n in nets]
        # the operand that leaves the prefix feeds each network's own (ungrouped) transposed convolution: row-padded layout
        a_next = rt.new_split(G * B, res, res, nxt[0].pack().Cin_pad, dev, C=nxt[0].in_channels,
                              pad_row=last and nxt[0].up == 2 and rt.pad_row_wanted(res, res), fmt=nxt[0].fmt())
        x32 = torch.empty((G * B, res, res, pack1.Cout), dtype=torch.float32, device=dev) if last else None
        nz, ns, gs = layer_noise(conv1s)
        rt.conv_same(a1.hi, a1.lo, pack1, pack1.Cin_pad, x32, dcoef=cat_layer(lo_i + i, 1), noise=nz, noise_strength=ns,
                     bias=_stack_cached(conv1s[0], '_ia_g_bias', [c.bias for c in conv1s]), act=conv1s[0].activation, gain=conv1s[0].act_gain,
                     clamp=conv1s[0].conv_clamp, mode=1, e1=(a_next, cat_layer(spans[index + 1][0], 0)), e2=(a_rgb, cat_layer(lo_i + i + 1, 0)),
                     group=group(gs))
        rawc = torch.empty((G * B, res, res, packr.Cout), dtype=torch.float32, device=dev)
        rt.conv_same(a_rgb.hi, a_rgb.lo, packr, packr.Cin_pad, rawc, mode=0, group=group(0))
        bias_rgb = _stack_cached(rgbs[0], '_ia_g_bias', [t.bias for t in rgbs], pad_to=packr.Cout)
        img = rt.torgb_finish(rawc, bias_rgb, rgbs[0].conv_clamp, img, group=(G, B))
        a = a_next
    out = []
    for g, n in enumerate(nets):
        sl = slice(g * B, (g + 1) * B)
        im = img[sl]
        if n.img_channels != im.shape[-1]:
            im = im[..., :n.img_channels].contiguous()
        out.append(dict(index=k_last, x32=x32[sl], img=im, a_next=rt.Split(a_next.hi[sl], a_next.lo[sl] if a_next.lo is not None else None, img_rows=a_next.img_rows),
                        styles=passes[g][0], dcoefs=passes[g][1], spans=passes[g][2]))
    return out


@persistence.persistent_class
class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_ws=-1, mapping_kwargs={},
                 **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        if mapping_ws == -1:
            mapping_ws = self.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=mapping_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)
