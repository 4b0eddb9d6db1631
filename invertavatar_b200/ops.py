"""Function-level mirror of the reference's ``torch_utils/ops`` Python API on the B200 engine: same names, argument
meaning and error behaviour (``bias_act.py:54``, ``upfirdn2d.py:72-350``, ``filtered_lrelu.py:58``,
``conv2d_resample.py:48``, ``fma.py:17``, ``conv2d_gradfix.py:37-45``, ``grid_sample_gradfix.py:28``), forward only.

Every function launches kernels of libinvertavatar_b200.so; CPU tensors raise RuntimeError (the reference falls back
to its ``_ref`` implementations there -- this build deliberately has no CPU path, see oracle/ for the checker)."""
import math

import numpy as np
import torch

from . import runtime as rt
from .persistence import EasyDict
from .stylegan2 import setup_filter  # noqa: F401  (re-exported: upfirdn2d.setup_filter)

# name -> def_alpha, def_gain, cuda_idx (bias_act.py:23-33); `func`/`ref` exist for API parity with code that introspects them
activation_funcs = {
    name: EasyDict(func=None, def_alpha=rt.ACT_DEFAULTS[name][0], def_gain=rt.ACT_DEFAULTS[name][1], cuda_idx=rt.ACT_IDS[name],
                   ref='', has_2nd_grad=name in ('tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'))
    for name in rt.ACT_IDS
}


def _check_impl(impl):
    assert impl in ['ref', 'cuda']


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    assert isinstance(x, torch.Tensor)
    _check_impl(impl)
    assert act in activation_funcs
    assert clamp is None or clamp >= 0
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
    return rt.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)


# ---- upfirdn2d.py ------------------------------------------------------------------------------------------------
def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, (int, np.integer)) for v in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = [int(v) for v in padding]
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    _check_impl(impl)
    if f is not None:
        assert isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32
    return rt.upfirdn2d(x, f, up=_parse_scaling(up), down=_parse_scaling(down), padding=_parse_padding(padding),
                        flip_filter=flip_filter, gain=float(gain))


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _parse_scaling(up)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


# ---- filtered_lrelu.py ---------------------------------------------------------------------------------------------
def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None, flip_filter=False,
                   impl='cuda'):
    """bias -> upsample FIR -> leaky ReLU * gain, clamp -> downsample FIR (filtered_lrelu.py:58-121).  impl='cuda': the
    library's ia_filtered_lrelu (two kernels through an fp32 workspace, the replacement of
    filtered_lrelu_plugin.filtered_lrelu, filtered_lrelu.cpp:20); impl='ref': the composition of the bias_act and upfirdn2d
    kernels -- the generic path the reference itself takes when its plugin answers rc = -1 (filtered_lrelu.py:225-231).
    Only StyleGAN3 calls this op and no inference script instantiates StyleGAN3."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    _check_impl(impl)
    fu_w, fu_h = _get_filter_size(fu)
    fd_w, fd_h = _get_filter_size(fd)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.dtype == x.dtype and b.shape[0] == x.shape[1]
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    px0, px1, py0, py1 = _parse_padding(padding)
    assert gain == float(gain) and gain > 0
    assert slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    B, Cc, in_h, in_w = x.shape
    in_dtype = x.dtype
    out_w = (in_w * up + (px0 + px1) - (fu_w - 1) - (fd_w - 1) + (down - 1)) // down
    out_h = (in_h * up + (py0 + py1) - (fu_h - 1) - (fd_h - 1) + (down - 1)) // down
    if impl == 'cuda' and x.dtype in (torch.float16, torch.float32):
        y = rt.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=[px0, px1, py0, py1], gain=gain, slope=slope, clamp=clamp,
                              flip_filter=flip_filter)
        assert tuple(y.shape) == (B, Cc, out_h, out_w) and y.dtype == in_dtype
        return y
    x = bias_act(x=x, b=b)
    x = upfirdn2d(x=x, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = bias_act(x=x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    x = upfirdn2d(x=x, f=fd, down=down, flip_filter=flip_filter)
    assert tuple(x.shape) == (B, Cc, out_h, out_w) and x.dtype == in_dtype
    return x


# ---- conv2d_resample.py / conv2d_gradfix.py ----------------------------------------------------------------------------
def _operand(x_nchw, Cin_pad):
    a, _ = rt.enc_prep([x_nchw.permute(0, 2, 3, 1)], C_pad=Cin_pad)
    return a


def _conv_same_plain(x, w, flip_weight=True):
    """stride-1 'same' correlation (flip_weight=True) / convolution (False) of NCHW x with [O,I,k,k] weights, k odd."""
    if not flip_weight and w.shape[-1] > 1:
        w = w.flip([2, 3])
    pack = rt.ConvPack(w.contiguous(), need_wsq=False)
    a = _operand(x, pack.Cin_pad)
    B, H, W, _ = a.hi.shape
    raw = torch.empty((B, H, W, pack.Cout), dtype=torch.float32, device=x.device)
    rt.conv_same(a.hi, a.lo, pack, pack.Cin_pad, raw, mode=0)
    return raw.permute(0, 3, 1, 2)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:48-143 for the configurations the generator reaches: (up=1, down=1, 'same' padding) and
    (up=2, 3x3 weights, padding=1, the [1,3,3,1] filter, flip_weight=False) -- the tensor-core transposed convolution +
    4x4 FIR of every ``conv0`` layer.  Anything else raises NotImplementedError rather than silently taking a slow path."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    out_c, in_c, kh, kw = [int(s) for s in w.shape]
    px0, px1, py0, py1 = _parse_padding(padding)
    if groups != 1 or down != 1 or kh != kw or kh % 2 == 0:
        raise NotImplementedError('conv2d_resample: only groups=1, down=1 and odd square kernels are on the generator path')
    if up == 1:
        if not (px0 == px1 == py0 == py1 == kh // 2):
            raise NotImplementedError("conv2d_resample(up=1): only 'same' padding (k//2) is supported")
        return _conv_same_plain(x.float(), w.float(), flip_weight=flip_weight).to(x.dtype)
    from .stylegan2 import _is_1331
    if up == 2 and kh == 3 and (px0, px1, py0, py1) == (1, 1, 1, 1) and not flip_weight and _is_1331(f):
        pack = rt.ConvPack(w.float().contiguous(), need_wsq=False)
        a = _operand(x.float(), pack.Cin_pad)
        B, H, W, _ = a.hi.shape
        raw = torch.empty((B, 2 * H + 1, 2 * W + 1, out_c), dtype=torch.float32, device=x.device)
        rt.conv_transpose_up2_raw(a.hi, a.lo, pack, pack.Cin_pad, raw)
        out = torch.empty((B, 2 * H, 2 * W, out_c), dtype=torch.float32, device=x.device)
        rt.fir_epilogue(raw, rt.fir4x4_gain4(x.device), out, None, None, None, None, 'linear', 1.0, None)
        return out.permute(0, 3, 1, 2).to(x.dtype)
    raise NotImplementedError('conv2d_resample: unsupported up/filter/padding combination (generator path: up=2, 3x3, pad 1, [1,3,3,1])')


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """conv2d_gradfix.conv2d forward for stride-1 'same' convolutions (the only call shape of the inference path)."""
    k = int(weight.shape[-1])
    pad = padding if isinstance(padding, int) else padding[0]
    if stride not in (1, (1, 1)) or dilation not in (1, (1, 1)) or groups != 1 or pad != k // 2 or k % 2 == 0:
        raise NotImplementedError("conv2d: only stride-1, groups=1, 'same'-padded convolutions are implemented on this engine")
    y = _conv_same_plain(input.float(), weight.float(), flip_weight=True)
    if bias is not None:
        y = rt.bias_act(y, bias.float(), dim=1)
    return y.to(input.dtype)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    raise NotImplementedError('conv_transpose2d: reached only through conv2d_resample(up=2) on this engine')


def fma(a, b, c):
    """a * b + c (fma.py:17); operands are broadcast, the result is computed by the bias_act kernel family's affine pass."""
    a, b, c = torch.broadcast_tensors(a, b, c)
    shape = a.shape
    n = a.numel()
    x = a.float().reshape(1, 1, 1, n)
    # y = x*scale + shift with per-"channel" scale/shift: one channel per element keeps this a single launch
    return rt.enc_affine_act(x, scale=b.float().reshape(n).contiguous(), shift=c.float().reshape(n).contiguous()).reshape(shape).to(a.dtype)


def grid_sample(input, grid):
    """F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False) (grid_sample_gradfix.py:28)."""
    y = rt.grid_sample_nhwc(rt.to_nhwc(input), grid)
    return y.permute(0, 3, 1, 2)


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """training/networks_stylegan2.py:34-91 (== training_avatar_texture/networks_stylegan2_new.py:34-91), function form.
    The hot path uses SynthesisLayer / ToRGBLayer (fused chains); this entry point serves external callers."""
    from .stylegan2 import _is_1331
    B, Cin = int(x.shape[0]), int(x.shape[1])
    O, I, kh, kw = [int(s) for s in weight.shape]
    assert I == Cin and tuple(styles.shape) == (B, Cin)
    if down != 1 or kh != kw or padding != kh // 2:
        raise NotImplementedError("modulated_conv2d: only down=1 with 'same' padding is on the generator path")
    if up == 2 and not (kh == 3 and not flip_weight and _is_1331(resample_filter)):
        raise NotImplementedError('modulated_conv2d(up=2): 3x3 weights, flip_weight=False and the [1,3,3,1] filter only')
    if up not in (1, 2) or (up == 1 and not flip_weight and kh > 1):
        raise NotImplementedError('modulated_conv2d: unsupported up / flip_weight combination')
    dev = x.device
    pack = rt.ConvPack(weight.float().contiguous(), need_wsq=demodulate)
    styles = styles.float().contiguous()
    dcoef = None
    if demodulate:   # d[b,o] = rsqrt(sum_i wsq[o,i] * s[b,i]^2 + 1e-8): one StylePlan row per sample (affine.weight = 0, bias = s[b])
        zeros = torch.zeros((Cin, 8), dtype=torch.float32, device=dev)
        wz = torch.zeros((1, 1, 8), dtype=torch.float32, device=dev)
        rows = []
        for b in range(B):
            plan = rt.StylePlan([dict(affine_w=zeros, affine_b=styles[b].contiguous(), wsq=pack.wsq, Cin=Cin, Cout=O, w_index=0, style_gain=1.0)], dev)
            _, dc = plan.run(wz)
            rows.append(dc[0])
        dcoef = torch.cat(rows, dim=0)
    hi, lo = rt.modsplit(rt.to_nhwc(x), styles, C_pad=pack.Cin_pad)
    H, W = int(x.shape[2]), int(x.shape[3])
    out = torch.empty((B, H * up, W * up, O), dtype=torch.float32, device=dev)
    nz = ns = None
    if noise is not None:
        nz = noise.float().expand(B, 1, H * up, W * up).reshape(B, H * up, W * up).contiguous() if noise.ndim == 4 else noise.float().contiguous()
        ns = torch.ones([], dtype=torch.float32, device=dev)
    if up == 1:
        rt.conv_same(hi, lo, pack, pack.Cin_pad, out, dcoef=dcoef, noise=nz, noise_strength=ns, act='linear', gain=1.0, mode=1)
    else:
        raw = torch.empty((B, 2 * H + 1, 2 * W + 1, O), dtype=torch.float32, device=dev)
        rt.conv_transpose_up2_raw(hi, lo, pack, pack.Cin_pad, raw)
        rt.fir_epilogue(raw, rt.fir4x4_gain4(dev), out, dcoef, nz, ns, None, 'linear', 1.0, None)
    return out.permute(0, 3, 1, 2).to(x.dtype)
