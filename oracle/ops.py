"""Oracle restatement of the reference's ``torch_utils.ops`` reference branches.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)
# name -> (default alpha, default gain); reference torch_utils/ops/bias_act.py:23-33
ACT_DEFAULTS = {
    'linear': (0.0, 1.0), 'relu': (0.0, SQRT2), 'lrelu': (0.2, SQRT2), 'tanh': (0.0, 1.0),
    'sigmoid': (0.0, 1.0), 'elu': (0.0, 1.0), 'selu': (0.0, 1.0), 'softplus': (0.0, 1.0),
    'swish': (0.0, SQRT2),
}


def _act(x, act, alpha):
    if act == 'linear':
        return x
    if act == 'relu':
        return F.relu(x)
    if act == 'lrelu':
        return F.leaky_relu(x, alpha)
    if act == 'tanh':
        return torch.tanh(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if act == 'elu':
        return F.elu(x)
    if act == 'selu':
        return F.selu(x)
    if act == 'softplus':
        return F.softplus(x)
    if act == 'swish':
        return torch.sigmoid(x) * x
    raise ValueError(act)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """y = clamp(act(x + b) * gain).  Follows ``_bias_act_ref``,
    reference torch_utils/ops/bias_act.py:93-122."""
    d_alpha, d_gain = ACT_DEFAULTS[act]
    alpha = float(d_alpha if alpha is None else alpha)
    gain = float(d_gain if gain is None else gain)
    if b is not None:
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    x = _act(x, act, alpha)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def setup_filter(taps=(1, 3, 3, 1), gain=1.0):
    """Normalised 2-D FIR from 1-D taps (outer product because < 8 taps).
    Reference torch_utils/ops/upfirdn2d.py:72-116."""
    f = torch.as_tensor(taps, dtype=torch.float32)
    if f.ndim == 1:
        f = torch.outer(f, f)
    f = f / f.sum()
    return f * (gain ** (f.ndim / 2))


def upfirdn2d(x, f, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """Zero-stuff by ``up``, pad/crop, correlate with the (flipped) filter, decimate by
    ``down``.  ``padding`` = [x0, x1, y0, y1].  Follows ``_upfirdn2d_ref``, reference
    torch_utils/ops/upfirdn2d.py:169-213 (2-D filter branch)."""
    B, C, H, W = x.shape
    px0, px1, py0, py1 = [int(p) for p in padding]
    if f is None:
        f = torch.ones(1, 1, dtype=torch.float32)
    z = x.new_zeros(B, C, H * up, W * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    k = f * (gain ** (f.ndim / 2))
    k = k.to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        z = F.conv2d(z, k[None, None].repeat(C, 1, 1, 1), groups=C)
    else:
        z = F.conv2d(z, k[None, None, None, :].repeat(C, 1, 1, 1), groups=C)
        z = F.conv2d(z, k[None, None, :, None].repeat(C, 1, 1, 1), groups=C)
    return z[:, :, ::down, ::down]


def upsample2d(x, f, up=2, gain=1.0):
    """Reference torch_utils/ops/upfirdn2d.py:315-350."""
    fw = f.shape[-1]
    fh = f.shape[0]
    p = [(fw + up - 1) // 2, (fw - up) // 2, (fh + up - 1) // 2, (fh - up) // 2]
    return upfirdn2d(x, f, up=up, padding=p, gain=gain * up * up)


def conv2d_resample(x, w, f=None, up=1, padding=0, groups=1, flip_weight=True):
    """The two branches the hot path takes: up=1 -> plain conv2d with symmetric padding;
    up=2 -> stride-2 transposed conv followed by the FIR.  Reference
    torch_utils/ops/conv2d_resample.py:48-143 (padding arithmetic :88-99,:114-131)."""
    kh, kw = w.shape[-2:]
    px0 = px1 = py0 = py1 = int(padding)
    if up == 1:
        wk = w if flip_weight or (kh == 1 and kw == 1) else w.flip([2, 3])
        return F.conv2d(x, wk, padding=[py0, px0], groups=groups)
    fw, fh = f.shape[-1], f.shape[0]
    px0 += (fw + up - 1) // 2
    px1 += (fw - up) // 2
    py0 += (fh + up - 1) // 2
    py1 += (fh - up) // 2
    out_ch = w.shape[0]
    if groups == 1:
        wt = w.transpose(0, 1)
    else:
        icg = w.shape[1]
        wt = w.reshape(groups, out_ch // groups, icg, kh, kw).transpose(1, 2)
        wt = wt.reshape(groups * icg, out_ch // groups, kh, kw)
    px0 -= kw - 1
    px1 -= kw - up
    py0 -= kh - 1
    py1 -= kh - up
    pxt = max(min(-px0, -px1), 0)
    pyt = max(min(-py0, -py1), 0)
    # _conv2d_wrapper(transpose=True, flip_weight=not flip_weight): flip when flip_weight is True
    wk = wt.flip([2, 3]) if (flip_weight and (kh > 1 or kw > 1)) else wt
    y = F.conv_transpose2d(x, wk, stride=up, padding=[pyt, pxt], groups=groups)
    return upfirdn2d(y, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2)


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=SQRT2, slope=0.2,
                   clamp=None, flip_filter=False):
    """bias -> upsample FIR -> lrelu*gain, clamp -> downsample FIR.  Follows
    ``_filtered_lrelu_ref``, reference torch_utils/ops/filtered_lrelu.py:123-159."""
    if isinstance(padding, int):
        padding = [padding] * 4
    px0, px1, py0, py1 = padding
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    x = upfirdn2d(x, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    x = upfirdn2d(x, fd, down=down, flip_filter=flip_filter)
    return x
