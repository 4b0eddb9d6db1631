"""Host-side launch layer: torch tensors in, C-ABI calls out.

PyTorch is used for device memory, streams and module/parameter bookkeeping only; every arithmetic step of
the hot path is a kernel in libinvertavatar_b200.so.  All functions require CUDA tensors and raise
RuntimeError otherwise -- there is no CPU path in the product (the CPU restatement lives in ``oracle/`` and
is test infrastructure)."""
import ctypes as C
import math
import threading

import numpy as np
import os
import torch

from . import _C

ACT_IDS = {'linear': 1, 'relu': 2, 'lrelu': 3, 'tanh': 4, 'sigmoid': 5, 'elu': 6, 'selu': 7, 'softplus': 8, 'swish': 9}
ACT_PRELU = 10      # IA_ACT_PRELU: convolution epilogues only (per-channel slopes); not one of the reference's bias_act activations
ACT_DEFAULTS = {  # name -> (def_alpha, def_gain), reference torch_utils/ops/bias_act.py:23-33
    'linear': (0.0, 1.0), 'relu': (0.0, math.sqrt(2)), 'lrelu': (0.2, math.sqrt(2)), 'tanh': (0.0, 1.0),
    'sigmoid': (0.0, 1.0), 'elu': (0.0, 1.0), 'selu': (0.0, 1.0), 'softplus': (0.0, 1.0), 'swish': (0.0, math.sqrt(2)),
}

_tls = threading.local()
_conv_impl = 'tc'   # 'tc' (tcgen05 tensor cores) | 'simt' (CUDA cores; bring-up / cross-check)


def set_conv_impl(name):
    global _conv_impl
    assert name in ('tc', 'simt')
    _conv_impl = name


def get_conv_impl():
    return _conv_impl


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('invertavatar_b200: CUDA tensors required (no CPU fallback exists for the hot path); '
                               f'got a tensor on {t.device}')


class _WrongDevice(Exception):
    """Raised by _enter when the tensor's device is not the thread's current CUDA device; _device_guarded re-runs the call
    under torch.cuda.device(idx) and restores the caller's device afterwards."""

    def __init__(self, idx):
        super().__init__(idx)
        self.idx = idx


def _enter(t):
    """Check that the tensor's device is the thread's current device (in torch and in the library's own runtime instance) and
    return the handle of torch's current stream on it.  The comparison is made on every call -- the current device may have
    been changed by anybody since the last one -- and a mismatch is resolved by the _device_guarded wrapper of the public
    entry points, which switches for the duration of the call and restores the caller's device (the OptionalCUDAGuard of
    the reference plugins, bias_act.cpp:58)."""
    _require_cuda(t)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if torch.cuda.current_device() != idx:
        raise _WrongDevice(idx)
    if getattr(_tls, 'device', None) != idx:
        _C.check(_C.lib().ia_set_device(idx), 'ia_set_device')
        _tls.device = idx
    return torch.cuda.current_stream(t.device).cuda_stream


def _device_guarded(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except _WrongDevice as e:
            prev = torch.cuda.current_device()
            try:
                with torch.cuda.device(e.idx):
                    return fn(*args, **kwargs)
            finally:
                # torch restored its own current device; bring the library's runtime instance back as well
                _C.check(_C.lib().ia_set_device(prev), 'ia_set_device')
                _tls.device = prev
    return wrapper


def _p(t):
    return None if t is None else t.data_ptr()


def _f32c(t):
    """fp32 + contiguous (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count():
    return int(_C.lib().ia_launch_count())


def reset_launch_count():
    _C.lib().ia_reset_launch_count()


def profile_begin():
    """Bracket every kernel launch of the library with CUDA events until profile_report()."""
    _C.check(_C.lib().ia_profile_begin(), 'ia_profile_begin')


def profile_report():
    """-> {entry point: {'ms': total device time, 'launches': n}} since profile_begin(); synchronises the device."""
    import json
    buf = C.create_string_buffer(1 << 16)
    n = _C.lib().ia_profile_report(buf, len(buf))
    if n < 0:
        _C.check(1, 'ia_profile_report')
    return json.loads(buf.value.decode())


# ---------------------------------------------------------------------------------------------------
# algorithmic-work counter (bench.py roofline leg): 2*MAC of every convolution launched between flop_count_begin() and
# flop_count_end(), counted as SURVEY 8(d) does -- true (unpadded) channel counts, transposed convolutions at their input
# resolution, strided encoder convolutions at their output resolution ('computed' is what the device really evaluates)
# ---------------------------------------------------------------------------------------------------
_flops = None


def flop_count_begin():
    global _flops
    _flops = {'algorithmic': 0, 'computed': 0, 'issued_mma': 0, 'launches': 0}


def flop_count_end():
    global _flops
    out, _flops = _flops, None
    return out


def _count_conv(B, H, W, pack, taps, stride=1, terms=3):
    if _flops is not None:
        f = 2 * B * H * W * pack.Cin * pack.Cout * taps
        _flops['computed'] += f
        _flops['algorithmic'] += f // (stride * stride)
        _flops['issued_mma'] += 2 * B * H * W * pack.Cin_pad * pack.Cout_pad * taps * terms
        _flops['launches'] += 1


# ---------------------------------------------------------------------------------------------------
# layout helpers: public tensors are logical NCHW with channels-last strides, kernels see NHWC
# ---------------------------------------------------------------------------------------------------
def to_nhwc(x):
    """[B,C,H,W] (any strides) -> contiguous [B,H,W,C] fp32 view/copy."""
    if x.dtype != torch.float32:
        x = x.float()
    return x.permute(0, 2, 3, 1).contiguous()


def from_nhwc(x):
    """contiguous [B,H,W,C] -> logical [B,C,H,W] view (channels_last strides, no copy)."""
    return x.permute(0, 3, 1, 2)


# ---------------------------------------------------------------------------------------------------
# torch_utils/ops equivalents
# ---------------------------------------------------------------------------------------------------
DTYPE_IDS = {torch.float32: 0, torch.float16: 1, torch.float64: 2}     # IA_DTYPE_* (the element types the reference plugins dispatch)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    _require_cuda(x, b)
    spec = ACT_DEFAULTS[act]
    alpha = float(spec[0] if alpha is None else alpha)
    gain = float(spec[1] if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if x.numel() == 0:
        return torch.empty_like(x)
    st = _enter(x)
    xd = x if x.dtype in DTYPE_IDS else x.float()      # f16 / f32 / f64 run natively (bias_act.cpp:81); anything else via fp32
    # dense in either contiguous or channels-last order: operate in memory order
    if xd.is_contiguous():
        xm, shape = xd, list(xd.shape)
        mdim = dim
    elif xd.ndim == 4 and xd.is_contiguous(memory_format=torch.channels_last):
        xm = xd.permute(0, 2, 3, 1)
        shape = list(xm.shape)
        mdim = {0: 0, 1: 3, 2: 1, 3: 2}[dim % 4]
    else:
        xd = xd.contiguous()
        xm, shape, mdim = xd, list(xd.shape), dim
    y = torch.empty_like(xd)
    Cn, inner = 1, 1
    if b is not None:
        if b.ndim != 1 or b.shape[0] != shape[mdim]:
            raise RuntimeError('bias_act: bias must be 1-D and match x.shape[dim]')
        b = b.to(xd.dtype).contiguous()
        Cn = shape[mdim]
        inner = int(np.prod(shape[mdim + 1:])) if mdim + 1 < len(shape) else 1
    _C.check(_C.lib().ia_bias_act(_p(xd), _p(b), _p(y), xd.numel(), Cn, inner, ACT_IDS[act], alpha, gain, clamp, DTYPE_IDS[xd.dtype], st), 'ia_bias_act')
    return y if y.dtype == x.dtype else y.to(x.dtype)


def upfirdn2d(x, f, up=(1, 1), down=(1, 1), padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """x [N,C,H,W] any strides; f fp32 [fh,fw] | [taps] (separable) | None.  padding = [x0,x1,y0,y1]."""
    _require_cuda(x, f)
    st = _enter(x)
    upx, upy = up
    downx, downy = down
    px0, px1, py0, py1 = [int(v) for v in padding]
    xd = x if x.dtype in DTYPE_IDS else x.float()      # f16 / f32 / f64 run natively (upfirdn2d.cpp:67)
    N, Cc, H, W = xd.shape
    if f is None:
        f = torch.ones(1, 1, dtype=torch.float32, device=x.device)
    f = f.to(torch.float32)
    passes = []
    if f.ndim == 2:
        passes.append((f.contiguous(), gain, (px0, px1, py0, py1), (upx, upy), (downx, downy)))
    else:  # separable: x pass then y pass, gain split as sqrt per pass (reference upfirdn2d.py:197,206-208)
        g = float(gain) ** 0.5
        passes.append((f.reshape(1, -1).contiguous(), g, (px0, px1, 0, 0), (upx, 1), (downx, 1)))
        passes.append((f.reshape(-1, 1).contiguous(), g, (0, 0, py0, py1), (1, upy), (1, downy)))
    cur = xd
    for (ff, g, (a0, a1, b0, b1), (ux, uy), (dx, dy)) in passes:
        n, c, h, w = cur.shape
        fh, fw = ff.shape
        outW = (w * ux + a0 + a1 - fw) // dx + 1
        outH = (h * uy + b0 + b1 - fh) // dy + 1
        if outW < 1 or outH < 1:
            raise RuntimeError('upfirdn2d: upsampled/padded signal is smaller than the filter')
        cl = cur.ndim == 4 and cur.stride(1) == 1 and c > 1
        y = torch.empty((n, c, outH, outW), dtype=cur.dtype, device=cur.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        p = _C.Upfirdn2dParams(_p(cur), _p(ff), _p(y), n, c, h, w, outH, outW, fh, fw, ux, uy, dx, dy, a0, b0,
                               1 if flip_filter else 0, float(g),
                               cur.stride(0), cur.stride(1), cur.stride(2), cur.stride(3),
                               y.stride(0), y.stride(1), y.stride(2), y.stride(3), DTYPE_IDS[cur.dtype])
        _C.check(_C.lib().ia_upfirdn2d(C.byref(p), st), 'ia_upfirdn2d')
        cur = y
    return cur if cur.dtype == x.dtype else cur.to(x.dtype)


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=(0, 0, 0, 0), gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    """filtered_lrelu_plugin.filtered_lrelu forward (filtered_lrelu.cpp:20): x [N,C,H,W] f16/f32 (any strides), fu/fd fp32 1-D
    (separable) / 2-D / None, b [C] or None, padding = [px0,px1,py0,py1] -> y [N,C,outH,outW] in x's dtype and memory format."""
    _require_cuda(x, fu, fd, b)
    st = _enter(x)
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError('filtered_lrelu: x and b must be float16 or float32')       # filtered_lrelu.cpp:33
    px0, px1, py0, py1 = [int(v) for v in padding]
    N, Cc, H, W = x.shape

    def fshape(f):
        if f is None:
            return None, 1, 1, 0
        f = f.to(torch.float32).contiguous()
        assert f.ndim in (1, 2)
        return f, int(f.shape[-1]), (int(f.shape[0]) if f.ndim == 2 else int(f.shape[-1])), (int(f.shape[0]) if f.ndim == 2 else 0)
    fu_t, fuw, fuh, fuh_field = fshape(fu)
    fd_t, fdw, fdh, fdh_field = fshape(fd)
    cw = W * up + px0 + px1 - (fuw - 1)
    ch = H * up + py0 + py1 - (fuh - 1)
    outW = (cw - (fdw - 1) + (down - 1)) // down
    outH = (ch - (fdh - 1) + (down - 1)) // down
    if outW < 1 or outH < 1:
        raise RuntimeError('filtered_lrelu: output must be at least 1x1')
    cl = x.stride(1) == 1 and Cc > 1
    y = torch.empty((N, Cc, outH, outW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last if cl else torch.contiguous_format)
    if b is not None:
        b = b.to(x.dtype).contiguous()
    p = _C.FilteredLreluParams()
    p.x, p.y, p.b, p.fu, p.fd = _p(x), _p(y), _p(b), _p(fu_t), _p(fd_t)
    p.N, p.C, p.inH, p.inW, p.outH, p.outW = N, Cc, H, W, outH, outW
    p.fuw, p.fuh, p.fdw, p.fdh = fuw, fuh_field, fdw, fdh_field
    p.up, p.down, p.px0, p.px1, p.py0, p.py1 = int(up), int(down), px0, px1, py0, py1
    p.gain, p.slope, p.clamp, p.flip = float(gain), float(slope), float(-1 if clamp is None else clamp), 1 if flip_filter else 0
    p.xs_n, p.xs_c, p.xs_h, p.xs_w = x.stride()
    p.ys_n, p.ys_c, p.ys_h, p.ys_w = y.stride()
    p.dtype = DTYPE_IDS[x.dtype]
    need = int(_C.lib().ia_filtered_lrelu_workspace(C.byref(p)))
    if need < 0:
        _C.check(1, 'ia_filtered_lrelu_workspace')
    ws = torch.empty(max(need, 4), dtype=torch.uint8, device=x.device)
    p.workspace, p.workspace_bytes = _p(ws), need
    _C.check(_C.lib().ia_filtered_lrelu(C.byref(p), st), 'ia_filtered_lrelu')
    return y


def filtered_lrelu_act_(x, gain=math.sqrt(2), slope=0.2, clamp=None):
    """In place clamp(lrelu(x) * gain) (filtered_lrelu_plugin.filtered_lrelu_act_, filtered_lrelu.cpp:217, without sign tensors)."""
    _require_cuda(x)
    st = _enter(x)
    if x.dtype not in DTYPE_IDS or not (x.is_contiguous() or (x.ndim == 4 and x.is_contiguous(memory_format=torch.channels_last))):
        raise RuntimeError('filtered_lrelu_act_: x must be a dense float16 / float32 / float64 tensor')
    _C.check(_C.lib().ia_filtered_lrelu_act(_p(x), x.numel(), DTYPE_IDS[x.dtype], None, 0, 0, float(gain), float(slope),
                                            float(-1 if clamp is None else clamp), 0, None, st), 'ia_filtered_lrelu_act')
    return x


# ---------------------------------------------------------------------------------------------------
# small dense layers
# ---------------------------------------------------------------------------------------------------
def fully_connected(x, weight, bias=None, w_gain=1.0, b_gain=1.0, act='linear', alpha=None, act_gain=None):
    _require_cuda(x, weight, bias)
    st = _enter(x)
    spec = ACT_DEFAULTS[act]
    alpha = float(spec[0] if alpha is None else alpha)
    act_gain = float(spec[1] if act_gain is None else act_gain)
    x2 = _f32c(x.reshape(-1, x.shape[-1]))
    w = _f32c(weight)
    b = _f32c(bias) if bias is not None else None
    B, In = x2.shape
    Out = w.shape[0]
    if w.shape[1] != In:
        raise RuntimeError(f'fully_connected: input features {In} do not match weight {tuple(w.shape)}')
    y = torch.empty((B, Out), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().ia_fully_connected(_p(x2), _p(w), _p(b), _p(y), B, In, Out, float(w_gain), float(b_gain),
                                         ACT_IDS[act], alpha, act_gain, In, Out, st), 'ia_fully_connected')
    return y.reshape(*x.shape[:-1], Out)


def normalize_2nd_moment(x, eps=1e-8):
    _require_cuda(x)
    st = _enter(x)
    x2 = _f32c(x)
    y = torch.empty_like(x2)
    _C.check(_C.lib().ia_normalize_2nd_moment(_p(x2), _p(y), x2.shape[0], x2.shape[1], float(eps), x2.shape[1], x2.shape[1], st),
             'ia_normalize_2nd_moment')
    return y


def broadcast_truncate(w, w_avg, num_ws, psi=1.0, cutoff=None):
    _require_cuda(w, w_avg)
    st = _enter(w)
    w = _f32c(w)
    B, D = w.shape
    ws = torch.empty((B, num_ws, D), dtype=torch.float32, device=w.device)
    cut = num_ws if cutoff is None else int(cutoff)
    wa = _f32c(w_avg) if w_avg is not None else None
    _C.check(_C.lib().ia_broadcast_truncate(_p(w), _p(wa), _p(ws), B, num_ws, D, float(psi), cut, st), 'ia_broadcast_truncate')
    return ws


# ---------------------------------------------------------------------------------------------------
# modulated convolution pieces
# ---------------------------------------------------------------------------------------------------
def _pad_to(v, m):
    return (v + m - 1) // m * m


def _none():
    return None


class _EngineCache:
    """Engine-side caches hung on modules (packed weights, style plans) are derived data keyed by the parameters' storage:
    they are dropped, not copied, when the owning module is deep-copied or pickled (copy.deepcopy(G), persistence pickles)."""

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_none, ())


# Operand formats of the tensor-core convolution (IA_OPFMT_* of the C-ABI)
FMT_BF16X3 = 0      # bf16 hi/lo pair, hi*hi + hi*lo + lo*hi: fp32-grade products, 3 MMAs per k-step
FMT_F16X1 = 1       # one fp16 tensor, 1 MMA per k-step
_FMT_DTYPE = {FMT_BF16X3: torch.bfloat16, FMT_F16X1: torch.float16}
_FMT_TERMS = {FMT_BF16X3: 3, FMT_F16X1: 1}


def layer_fmt(layer):
    """Operand format a convolution layer runs in.  A layer carries ``tc_fmt`` when its owner measured that the cheaper
    format keeps the owner's output inside the parity budget (TriPlaneGenerator tags the 3x3 layers of its three backbones:
    single-pass fp16 there moves the final image by 1e-4..3e-4 max-abs, profiles/r2_conv_precision_probe*.json; the
    super-resolution layers and every ToRGB stay 3-term).  IA_CONV_PRECISION=bf16x3 forces the 3-term split everywhere
    (strict mode: every stage within 2e-4 of fp32)."""
    if os.environ.get('IA_CONV_PRECISION', 'auto') == 'bf16x3':
        return FMT_BF16X3
    return int(getattr(layer, 'tc_fmt', FMT_BF16X3))


class ConvPack(_EngineCache):
    """GEMM-layout weights of one conv layer: [taps][Cout_pad][Cin_pad] in operand format ``fmt`` (bf16 hi/lo, or fp16 in
    w_hi alone with w_lo = None) and wsq[Cout][Cin]."""

    def __init__(self, weight, need_wsq=True, fmt=FMT_BF16X3):
        _require_cuda(weight)
        st = _enter(weight)
        w = _f32c(weight.detach())
        self.Cout, self.Cin, self.kh, self.kw = w.shape
        self.taps = self.kh * self.kw
        self.Cout_pad = _pad_to(self.Cout, 32)
        self.Cin_pad = _pad_to(self.Cin, 64)
        self.fmt = int(fmt)
        dev = w.device
        self.w_hi = torch.empty((self.taps, self.Cout_pad, self.Cin_pad), dtype=_FMT_DTYPE[self.fmt], device=dev)
        self.w_lo = torch.empty_like(self.w_hi) if self.fmt == FMT_BF16X3 else None
        self.wsq = torch.empty((self.Cout, self.Cin), dtype=torch.float32, device=dev) if need_wsq else None
        _C.check(_C.lib().ia_pack_conv_weight(_p(w), self.Cout, self.Cin, self.kh, self.kw, self.Cout_pad, self.Cin_pad,
                                              _p(self.w_hi), _p(self.w_lo), _p(self.wsq), self.fmt, st), 'ia_pack_conv_weight')
        self.key = (weight.data_ptr(), weight._version, str(dev), self.fmt)

    @classmethod
    def from_tensors(cls, weight, fmt, w_hi, w_lo, wsq):
        """A pack rebuilt from tensors produced earlier by this class for the same weight values (the on-disk pack cache)."""
        self = cls.__new__(cls)
        self.Cout, self.Cin, self.kh, self.kw = [int(v) for v in weight.shape]
        self.taps = self.kh * self.kw
        self.Cout_pad, self.Cin_pad, self.fmt = _pad_to(self.Cout, 32), _pad_to(self.Cin, 64), int(fmt)
        assert tuple(w_hi.shape) == (self.taps, self.Cout_pad, self.Cin_pad) and w_hi.dtype == _FMT_DTYPE[self.fmt]
        self.w_hi, self.w_lo, self.wsq = w_hi, w_lo, wsq
        self.key = (weight.data_ptr(), weight._version, str(weight.device), self.fmt)
        return self

    @staticmethod
    def current(cache_owner, attr, weight, need_wsq=True, fmt=FMT_BF16X3):
        """Return the pack cached on ``cache_owner.<attr>``; repack when the parameter (or the requested format) changed."""
        pack = cache_owner.__dict__.get(attr)
        key = (weight.data_ptr(), weight._version, str(weight.device), int(fmt))
        if pack is None or pack.key != key:
            pack = ConvPack(weight, need_wsq, fmt)
            cache_owner.__dict__[attr] = pack
        return pack


def _packable(module):
    """(qualified name, module, weight, packer) of every convolution layer of ``module`` that owns a tensor-core pack."""
    for name, m in module.named_modules():
        w = getattr(m, 'weight', None)
        if not (isinstance(w, torch.Tensor) and w.is_cuda and w.ndim == 4):
            continue
        fn = getattr(m, 'pack', None)
        if callable(fn):
            yield name, m, w, fn
        elif isinstance(m, torch.nn.Conv2d) and m.groups == 1:      # encoder convolutions (plain nn.Conv2d parameters)
            yield name, m, w, (lambda m=m, w=w: ConvPack.current(m, '_ia_pack', w, need_wsq=False))


def pack_cache_path(source_hash, cache_dir=None):
    """File of the on-disk pack cache for a checkpoint: keyed by the checkpoint's content hash (legacy.load_network_pkl records
    the sha256 of the pickle bytes), the library's source hash (a new kernel layout must not read old packs) and the precision
    policy in force."""
    import hashlib
    from . import build as _build
    cache_dir = cache_dir or os.environ.get('IA_PACK_CACHE') or os.path.join(os.path.expanduser('~'), '.cache', 'invertavatar_b200', 'packs')
    tag = hashlib.sha256(('|'.join([str(source_hash), _build.source_hash(), os.environ.get('IA_CONV_PRECISION', 'auto'),
                                    str(_C.ABI_VERSION)])).encode()).hexdigest()[:32]
    return os.path.join(cache_dir, tag + '.pt')


def prepack(module, source_hash=None, cache_dir=None):
    """One-time weight packing at load: build the GEMM-layout weights (bf16 hi/lo or fp16, per the layer's
    precision) + the demodulation table sum w^2 of every convolution / ToRGB layer of ``module`` now instead of at first use.
    Call after ``module.to('cuda')`` (e.g. right after ``legacy.load_network_pkl``); returns the packed bytes.  Packs are keyed
    on the parameter's storage and version, so a later in-place update (PTI fine-tuning, ``copy_params_and_buffers``) repacks
    that layer transparently.

    ``source_hash`` (default: the hash ``legacy.load_network_pkl`` recorded on the module, if any) turns on the on-disk cache:
    the packs are written to ``pack_cache_path(source_hash)`` the first time and read back -- no packing kernels, no fp32 ->
    operand conversion -- on every later load of the same checkpoint with the same library."""
    source_hash = source_hash or module.__dict__.get('_ia_source_hash')
    path = pack_cache_path(source_hash, cache_dir) if source_hash else None
    layers = list(_packable(module))
    total = 0
    if path and os.path.exists(path):
        try:
            blob = torch.load(path, map_location=layers[0][2].device if layers else 'cpu', weights_only=True)
        except Exception:
            blob = None
        if blob is not None and set(blob.keys()) == {n for n, *_ in layers}:
            ok = True
            for name, m, w, fn in layers:
                e = blob[name]
                fmt = layer_fmt(m) if hasattr(m, 'fmt') else FMT_BF16X3
                if int(e['fmt']) != fmt or tuple(e['shape']) != tuple(w.shape):
                    ok = False
                    break
            if ok:
                for name, m, w, fn in layers:
                    e = blob[name]
                    m.__dict__['_ia_pack'] = ConvPack.from_tensors(w, int(e['fmt']), e['w_hi'], e.get('w_lo'), e.get('wsq'))
                    pk = m.__dict__['_ia_pack']
                    total += pk.w_hi.numel() * (4 if pk.w_lo is not None else 2) + (pk.wsq.numel() * 4 if pk.wsq is not None else 0)
                return total
    for name, m, w, fn in layers:
        pk = fn()
        total += pk.w_hi.numel() * (4 if pk.w_lo is not None else 2) + (pk.wsq.numel() * 4 if pk.wsq is not None else 0)
    if path:
        blob = {}
        for name, m, w, fn in layers:
            pk = m.__dict__['_ia_pack']
            e = {'fmt': torch.tensor(pk.fmt), 'shape': torch.tensor(list(w.shape)), 'w_hi': pk.w_hi}
            if pk.w_lo is not None:
                e['w_lo'] = pk.w_lo
            if pk.wsq is not None:
                e['wsq'] = pk.wsq
            blob[name] = e
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + f'.tmp{os.getpid()}'
        torch.save(blob, tmp)
        os.replace(tmp, path)
    return total


class ConvPackGroup(_EngineCache):
    """Packed weights of the same layer of several networks, stacked along the tap axis ([G*taps][Cout_pad][Cin_pad]) for a
    grouped launch (ia_conv_params.groups).  Output channels are zero-padded to the widest member (ToRGB layers of the
    32- and 96-channel backbones)."""

    def __init__(self, weights, need_wsq=True, fmt=FMT_BF16X3):
        Cout = max(int(w.shape[0]) for w in weights)
        packs = []
        for w in weights:
            w = w.detach()
            if w.shape[0] != Cout:
                wp = torch.zeros((Cout,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
                wp[:w.shape[0]] = w
                w = wp
            packs.append(ConvPack(w, need_wsq, fmt))
        p0 = packs[0]
        self.G = len(packs)
        self.fmt = int(fmt)
        self.Cout, self.Cin, self.kh, self.kw, self.taps = Cout, p0.Cin, p0.kh, p0.kw, p0.taps
        self.Cout_pad, self.Cin_pad = p0.Cout_pad, p0.Cin_pad
        self.w_hi = torch.cat([q.w_hi for q in packs], dim=0).contiguous()
        self.w_lo = torch.cat([q.w_lo for q in packs], dim=0).contiguous() if p0.w_lo is not None else None
        self.wsq = [q.wsq for q in packs]
        self.key = tuple((w.data_ptr(), w._version, str(w.device)) for w in weights) + (self.fmt,)

    @staticmethod
    def current(cache_owner, attr, weights, need_wsq=True, fmt=FMT_BF16X3):
        pack = cache_owner.__dict__.get(attr)
        key = tuple((w.data_ptr(), w._version, str(w.device)) for w in weights) + (int(fmt),)
        if pack is None or pack.key != key:
            pack = ConvPackGroup(weights, need_wsq, fmt)
            cache_owner.__dict__[attr] = pack
        return pack


class StylePlan(_EngineCache):
    """Device table of ia_style_layer entries + output buffers for a group of layers that share one ws tensor."""

    def __init__(self, entries, device, share=None):
        # entries: list of dict(affine_w, affine_b, wsq|None, Cin, Cout, w_index, style_gain)
        # share: optional list parallel to entries of None | (key, g, G): the entry's style / dcoef rows live in rows [g*B, (g+1)*B)
        # of a [G*B, C] buffer shared by the G entries with the same key (the same layer of G networks evaluated as one grouped
        # batch: the group-major vectors the grouped convolutions read exist without any concatenation)
        self.entries = entries
        self.device = device
        self.share = share
        self.shared = {}          # B -> {key: (styles [G*B,Cin], dcoefs [G*B,Cout] | None)}
        self._by_batch = {}       # B -> (styles, dcoefs, host table, device table): callers alternate between batch sizes (an identity of
                                  # eval_seq.py runs the backbones at B = 1 and the T-frame render at B = T), and a rebuild is an
                                  # unpinned H2D copy -- not allowed inside a CUDA-graph capture and a stall outside of one

    def _key(self, B):
        """Output buffers are static per batch size -- and per capture scope: a graph captured by graphs.GraphedCall gets its own,
        so that two graphs over the same module may replay concurrently (see _scratch_key)."""
        return (B, _active_scope)

    def shared_for(self, B):
        return self.shared[self._key(B)]

    def _build(self, B):
        n = len(self.entries)
        styles, dcoefs, shared = [], [], {}
        for i, e in enumerate(self.entries):
            sh = self.share[i] if self.share is not None else None
            if sh is None:
                styles.append(torch.empty((B, e['Cin']), dtype=torch.float32, device=self.device))
                dcoefs.append(torch.empty((B, e['Cout']), dtype=torch.float32, device=self.device) if e['wsq'] is not None else None)
                continue
            key, g, G = sh
            if key not in shared:
                shared[key] = (torch.empty((G * B, e['Cin']), dtype=torch.float32, device=self.device),
                               torch.empty((G * B, e['Cout']), dtype=torch.float32, device=self.device) if e['wsq'] is not None else None)
            S, D = shared[key]
            assert S.shape[1] == e['Cin'] and (D is None) == (e['wsq'] is None) and (D is None or D.shape[1] == e['Cout'])
            styles.append(S[g * B:(g + 1) * B])
            dcoefs.append(None if D is None else D[g * B:(g + 1) * B])
        self.shared[self._key(B)] = shared
        arr = (_C.StyleLayer * n)()
        for i, e in enumerate(self.entries):
            w_dim = e['affine_w'].shape[1]
            arr[i] = _C.StyleLayer(_p(e['affine_w']), _p(e['affine_b']), _p(e['wsq']), _p(styles[i]), _p(dcoefs[i]),
                                   e['Cin'], e['Cout'], e['w_index'], w_dim, 1.0 / math.sqrt(w_dim), float(e['style_gain']))
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        dev = torch.from_numpy(raw).to(self.device)
        built = (styles, dcoefs, arr, dev)
        self._by_batch[self._key(B)] = built
        return built

    def run(self, ws):
        """ws: [B, n, w_dim] fp32 with unit inner stride (may be a narrow() view of a wider tensor)."""
        st = _enter(ws)
        if ws.dtype != torch.float32 or ws.stride(2) != 1 or ws.stride(1) != ws.shape[2] or ws.stride(0) % ws.shape[2] != 0 or \
                ws.stride(0) < ws.shape[1] * ws.shape[2]:   # (an expand()ed batch has stride 0: materialise it)
            ws = ws.float().contiguous()
        B = ws.shape[0]
        built = self._by_batch.get(self._key(B)) or self._build(B)
        styles, dcoefs, host, dev = built
        num_ws = max(ws.stride(0) // ws.shape[2], ws.shape[1])  # batch stride of a narrow() view, in rows
        _C.check(_C.lib().ia_styles(_p(dev), host, len(self.entries), _p(ws), B, num_ws, st), 'ia_styles')
        return styles, dcoefs


def modsplit(x_nhwc, styles=None, cond=None, cond_alpha=None, C_pad=None, fmt=FMT_BF16X3):
    """x [B,H,W,C] fp32 (pixel stride = x.stride(2)) -> (hi, lo) [B,H,W,C_pad] of x*styles (after optional blend) in operand
    format ``fmt`` (lo is None for FMT_F16X1)."""
    sp = modsplit_split(x_nhwc, styles, cond, cond_alpha, C_pad, fmt=fmt)
    return sp.hi, sp.lo


def modsplit_split(x_nhwc, styles=None, cond=None, cond_alpha=None, C_pad=None, pad_row=False, fmt=FMT_BF16X3):
    """modsplit into a Split; pad_row: the row-padded layout of new_split(pad_row=True)."""
    st = _enter(x_nhwc)
    B, H, W, Cc = x_nhwc.shape
    assert x_nhwc.stride(3) == 1 and x_nhwc.stride(1) == W * x_nhwc.stride(2) and x_nhwc.stride(0) == H * x_nhwc.stride(1)
    C_pad = _pad_to(Cc, 64) if C_pad is None else C_pad
    sp = new_split(B, H, W, C_pad, x_nhwc.device, pad_row=pad_row, fmt=fmt)
    cond_ld = 0
    if cond is not None:
        assert cond.shape == x_nhwc.shape and cond.stride(3) == 1 and cond_alpha is not None
        assert cond.stride(1) == W * cond.stride(2) and cond.stride(0) == H * cond.stride(1)
        assert cond_alpha.is_contiguous() and cond_alpha.numel() == B * H * W
        cond_ld = cond.stride(2)
    p = _C.ModsplitParams(_p(x_nhwc), x_nhwc.stride(2), _p(styles), _p(cond), cond_ld, _p(cond_alpha), _p(sp.hi), _p(sp.lo),
                          B, H * W, Cc, C_pad, sp.img_pix, sp.fmt)
    _C.check(_C.lib().ia_modsplit(C.byref(p), st), 'ia_modsplit')
    return sp


class Split:
    """A tensor-core A operand: bf16 hi/lo pair [B,H,W,C_pad] holding x * styles of the consuming layer.  ``img_rows`` > H: the
    images are ``img_rows`` rows apart in memory and the extra rows are zero (hi / lo are views of the padded buffers) -- the
    layout a stride-2 transposed convolution tiles across image boundaries (ia_conv_params.a_img_rows)."""
    __slots__ = ('hi', 'lo', 'C_pad', 'img_rows', 'fmt')

    def __init__(self, hi, lo, img_rows=None):
        self.hi, self.lo, self.C_pad = hi, lo, hi.shape[-1]
        self.img_rows = int(hi.shape[1] if img_rows is None else img_rows)
        self.fmt = FMT_F16X1 if hi.dtype == torch.float16 else FMT_BF16X3       # (FMT_F16X1: lo is None)

    @property
    def img_pix(self):
        """Pixel stride between images (ia_emit.e1_img_pix / ia_modsplit_params.out_img_pix); 0 = dense."""
        return 0 if self.img_rows == self.hi.shape[1] else self.img_rows * self.hi.shape[2]


def pad_row_wanted(H, W, impl=None):
    """Whether the operand of a stride-2 transposed convolution with H x W input should carry a zero row after every image:
    tensor-core path and both the consumer's phase grids and the producer are handled by the persistent kernel."""
    return ((impl or _conv_impl) == 'tc' and H >= 32 and W >= 32 and os.environ.get('IA_CONV_CAT_ROWS', '1') != '0')


def new_split(B, H, W, C_pad, device, C=None, pad_row=False, fmt=FMT_BF16X3):
    """Uninitialised operand buffers in format ``fmt``; zero-filled when the producer leaves padding channels (C < C_pad)
    untouched.  pad_row: [B, H+1, W, C_pad] buffers with row H zeroed, exposed as [B, H, W, C_pad] views (see Split)."""
    make = torch.zeros if (C is not None and C != C_pad) else torch.empty
    rows = H + 1 if pad_row else H
    two = fmt == FMT_BF16X3
    hi = make((B, rows, W, C_pad), dtype=_FMT_DTYPE[fmt], device=device)
    lo = make((B, rows, W, C_pad), dtype=_FMT_DTYPE[fmt], device=device) if two else None
    if pad_row:
        if make is torch.empty:
            hi[:, H].zero_()
            if two:
                lo[:, H].zero_()
        return Split(hi[:, :H], lo[:, :H] if two else None, img_rows=rows)
    return Split(hi, lo)


def _emit(out32=None, e1=None, e2=None, rgb=None):
    """out32: fp32 NHWC tensor or None; e1/e2: (Split, styles [B,Cout] or None) pairs filled by the epilogue with
    split(v * styles) -- the operand of the next convolution / the ToRGB layer.  rgb = (raw [B,H,W,n] zero-filled fp32,
    weight [n,Cout] fp32, styles [B,Cout]): the ToRGB contraction itself, accumulated by the epilogue (n <= 4)."""
    e = _C.Emit()
    if rgb is not None:
        raw, w2d, st = rgb
        assert raw.is_contiguous() and w2d.is_contiguous() and st.is_contiguous() and w2d.shape[0] == raw.shape[-1] <= 4
        e.rgb_out, e.rgb_w, e.rgb_s, e.rgb_n = _p(raw), _p(w2d), _p(st), int(raw.shape[-1])
    if out32 is not None:
        e.out32 = _p(out32)
        e.out32_ld = out32.stride(2)
    if e1 is not None:
        sp, st = e1
        e.hi1, e.lo1, e.s1, e.c1_pad = _p(sp.hi), _p(sp.lo), _p(st), sp.C_pad
        e.e1_img_pix = sp.img_pix
        e.fmt1 = sp.fmt
    if e2 is not None:
        sp, st = e2
        e.hi2, e.lo2, e.s2, e.c2_pad = _p(sp.hi), _p(sp.lo), _p(st), sp.C_pad
        e.fmt2 = sp.fmt
    return e


# Persistent device state (split-K workspace + tickets, squeeze-excite sums, the static style / demodulation buffers of a StylePlan)
# is owned by ONE stream at a time.  Two CUDA graphs captured separately use the same capture-stream handle and the same modules,
# so keyed by (device, stream) / batch size alone they would share that state -- fine while their replays are serialised, a race
# (corrupted tickets, illegal addresses, another identity's styles) when they replay concurrently on two streams.
# graphs.GraphedCall / GraphedSynthesis therefore run their warm-up and capture inside ``capture_scope()``: state requested while a
# scope is active is keyed by the scope as well, i.e. private to that graph (built during the warm-up, found again by the capture).
_capture_scope_next = 0
_active_scope = 0


class capture_scope:
    def __enter__(self):
        global _capture_scope_next, _active_scope
        _capture_scope_next += 1
        self.prev, _active_scope = _active_scope, _capture_scope_next
        return _active_scope

    def __exit__(self, *exc):
        global _active_scope
        _active_scope = self.prev


def _scratch_key(device, st):
    return (device.index, int(st or 0), _active_scope)


_SPLITK = {}
_SPLITK_WS_BYTES = 48 << 20          # fp32 partial accumulators: tiles x 256 pixels x N tile x splits (<= ~150 tile-splits of 128 KB)
_SPLITK_COUNTERS = 4096              # 8 tickets per tile; split launches have fewer tiles than SMs


def _splitk_scratch(p, device, st):
    """Split-K workspace + ticket counters of (device, stream): owned by one stream at a time, as ia_conv_params asks (launches of
    a stream are ordered; the counters are zero between launches).  Persistent, so CUDA-graph captures may hold them."""
    if os.environ.get('IA_CONV_SPLITK', '') == '0':
        return
    key = _scratch_key(device, st)
    buf = _SPLITK.get(key)
    if buf is None:
        buf = (torch.empty(_SPLITK_WS_BYTES // 4, dtype=torch.float32, device=device), torch.zeros(_SPLITK_COUNTERS, dtype=torch.int32, device=device))
        _SPLITK[key] = buf
    p.splitk_ws, p.splitk_ws_bytes, p.splitk_counters, p.splitk_n_counters = _p(buf[0]), _SPLITK_WS_BYTES, _p(buf[1]), _SPLITK_COUNTERS


def _check_fmt(hi, lo, pack):
    """The activation operand must be in the format the weights were packed in."""
    fmt = FMT_F16X1 if hi.dtype == torch.float16 else FMT_BF16X3
    if fmt != pack.fmt or (fmt == FMT_BF16X3 and lo is None):
        raise RuntimeError(f'convolution operands disagree: activations in format {fmt}, weights packed in format {pack.fmt}')
    return fmt


def _conv_call(p, st, impl=None):
    impl = impl or _conv_impl
    fn = _C.lib().ia_conv_tc if impl == 'tc' else _C.lib().ia_conv_simt
    _C.check(fn(C.byref(p), st), 'ia_conv_' + impl)


def _noise_bstride(noise):
    return 0 if noise is None or noise.ndim < 3 else noise.shape[-1] * noise.shape[-2]


def _set_group(p, group):
    """group = (G, imgs_per_group, noise_gstride) or None."""
    if group is not None:
        p.groups, p.imgs_per_group, p.noise_gstride = int(group[0]), int(group[1]), int(group[2])


def conv_same(hi, lo, pack, Cin_pad, out32, dcoef=None, noise=None, noise_strength=None, bias=None, act='linear',
              gain=1.0, clamp=None, mode=1, impl=None, e1=None, e2=None, group=None, rgb=None, img_prev=None, alg_stride=1, slope=None,
              alpha=None):
    """k x k correlation, stride 1, 'same' padding (flip_weight=True branch of conv2d_resample, :134-136)."""
    st = _enter(hi)
    B, H, W, _ = hi.shape
    k = pack.kh
    pad = k // 2
    p = _C.ConvParams()
    p.a_hi, p.a_lo, p.B, p.H, p.W, p.Cin_pad = _p(hi), _p(lo), B, H, W, Cin_pad
    p.w_hi, p.w_lo, p.Cout, p.Cout_pad, p.n_taps_total = _p(pack.w_hi), _p(pack.w_lo), pack.Cout, pack.Cout_pad, pack.taps
    p.op_fmt = _check_fmt(hi, lo, pack)
    p.GH, p.GW, p.ntaps = H, W, k * k
    t = 0
    for ky in range(k):
        for kx in range(k):
            p.dy[t], p.dx[t], p.wtap[t] = ky - pad, kx - pad, ky * k + kx
            t += 1
    p.OH, p.OW, p.sy, p.sx, p.py, p.px = H, W, 1, 1, 0, 0
    p.mode = mode
    p.dcoef, p.noise, p.noise_strength, p.bias = _p(dcoef), _p(noise), _p(noise_strength), _p(bias)
    p.noise_bstride = _noise_bstride(noise)
    if act == 'prelu':
        assert slope is not None and slope.numel() == pack.Cout and slope.is_contiguous()
        p.act, p.alpha, p.slope = ACT_PRELU, 0.0, _p(slope)
    else:
        p.act, p.alpha = ACT_IDS[act], ACT_DEFAULTS[act][0] if alpha is None else float(alpha)
    p.gain, p.clamp = float(gain), float(-1 if clamp is None else clamp)
    p.emit = _emit(out32, e1, e2, rgb)
    if img_prev is not None:
        assert mode == 2 and img_prev.is_contiguous() and tuple(img_prev.shape) == (B, H // 2, W // 2, pack.Cout), (img_prev.shape, hi.shape)
        p.img_prev = _p(img_prev)
    _set_group(p, group)
    _count_conv(B, H, W, pack, k * k, stride=alg_stride, terms=_FMT_TERMS[pack.fmt])
    _splitk_scratch(p, hi.device, st)
    _conv_call(p, st, impl)


def can_fuse_torgb_tail(H, W, Cout, impl=None):
    """Whether conv_same(..., mode=2) -- ToRGB tail (bias, clamp, + upsampled previous image) inside the 1x1 convolution's
    epilogue -- is available: tensor-core path, persistent kernel, Cout % 4 == 0."""
    return ((impl or _conv_impl) == 'tc' and H * W >= 128 and W >= 8 and Cout % 4 == 0 and H % 2 == 0 and W % 2 == 0
            and os.environ.get('IA_FUSE_TORGB_TAIL', '1') != '0')


def can_fuse_torgb(H, W, Cout, n_img, impl=None):
    """Whether conv_same(..., rgb=...) is available for this layer: tensor-core path, persistent kernel (>= 128 pixels per image,
    width >= 8), <= 4 image channels, and at most two N tiles of 128 (two partial sums commute: the result stays exact)."""
    return ((impl or _conv_impl) == 'tc' and H * W >= 128 and W >= 8 and n_img <= 4 and Cout % 4 == 0 and Cout <= 256
            and os.environ.get('IA_FUSE_TORGB', '1') != '0')


def conv_transpose_up2_raw(hi, lo, pack, Cin_pad, raw, impl=None, group=None, img_rows=0):
    """Stride-2 transposed 3x3 convolution (true convolution, conv2d_resample.py:114-127) written as four output-parity
    phases; raw is [B, 2H+1, 2W+1, Cout] fp32.  Even output row 2m gets ky=0 from input row m and ky=2 from row m-1; odd
    output row 2m+1 gets ky=1 from row m (same for columns)."""
    st = _enter(hi)
    B, H, W, _ = hi.shape
    assert pack.kh == 3 and pack.kw == 3
    rowtaps = {0: [(0, 0), (2, -1)], 1: [(1, 0)]}
    phases = (_C.ConvParams * 4)()
    k = 0
    for py in (0, 1):
        for px in (0, 1):
            p = phases[k]
            k += 1
            p.a_hi, p.a_lo, p.B, p.H, p.W, p.Cin_pad = _p(hi), _p(lo), B, H, W, Cin_pad
            p.w_hi, p.w_lo, p.Cout, p.Cout_pad, p.n_taps_total = _p(pack.w_hi), _p(pack.w_lo), pack.Cout, pack.Cout_pad, pack.taps
            p.op_fmt = _check_fmt(hi, lo, pack)
            p.GH, p.GW = H + 1 - py, W + 1 - px
            t = 0
            for (ky, dy) in rowtaps[py]:
                for (kx, dx) in rowtaps[px]:
                    p.dy[t], p.dx[t], p.wtap[t] = dy, dx, ky * 3 + kx
                    t += 1
            p.ntaps = t
            p.OH, p.OW, p.sy, p.sx, p.py, p.px = 2 * H + 1, 2 * W + 1, 2, 2, py, px
            p.mode = 0
            p.act, p.alpha, p.gain, p.clamp = 1, 0.0, 1.0, -1.0
            p.emit = _emit(raw)
            p.a_img_rows = int(img_rows) if img_rows and img_rows != H else 0
            _set_group(p, group)
            _splitk_scratch(p, hi.device, st)
    _count_conv(B, H, W, pack, 9, terms=_FMT_TERMS[pack.fmt])
    if (impl or _conv_impl) == 'tc':
        _C.check(_C.lib().ia_conv_tc_phases(phases, 4, st), 'ia_conv_tc_phases')     # one persistent launch for the four phases
    else:
        for k in range(4):
            _conv_call(phases[k], st, impl)


_FIR_CACHE = {}


def fir4x4_gain4(device):
    """outer([1,3,3,1])/64 * 4 : the resample filter of every synthesis layer times the up**2 gain."""
    key = str(device)
    f = _FIR_CACHE.get(key)
    if f is None:
        k = torch.tensor([1.0, 3.0, 3.0, 1.0])
        f2 = torch.outer(k, k)
        f2 = f2 / f2.sum() * 4.0
        f = f2.to(device)
        _FIR_CACHE[key] = f
    return f


def fir_epilogue(raw, fir, out32, dcoef, noise, noise_strength, bias, act, gain, clamp, e1=None, e2=None, group=None):
    assert e1 is None or e1[0].img_pix == 0, 'fir_epilogue writes dense operands only'
    st = _enter(raw)
    B, RH, RW, Cc = raw.shape
    OH, OW = RH - 1, RW - 1
    p = _C.FirParams()
    p.raw, p.B, p.RH, p.RW, p.C = _p(raw), B, RH, RW, Cc
    p.fir, p.OH, p.OW = _p(fir), OH, OW
    p.dcoef, p.noise, p.noise_strength, p.bias = _p(dcoef), _p(noise), _p(noise_strength), _p(bias)
    p.noise_bstride = _noise_bstride(noise)
    p.act, p.alpha, p.gain, p.clamp = ACT_IDS[act], ACT_DEFAULTS[act][0], float(gain), float(-1 if clamp is None else clamp)
    p.emit = _emit(out32, e1, e2)
    _set_group(p, group)
    _C.check(_C.lib().ia_fir_epilogue(C.byref(p), st), 'ia_fir_epilogue')


class frame_sink:
    """Context manager: while active (on THIS host thread), the ToRGB tail that writes the final planar image (out_nchw) also
    stores every value into this rank's slot of every rank's gathered frame buffer (peer-mapped / multicast symmetric
    memory).  ``capacity`` = elements of the slot: an image batch of any other size raises instead of writing
    past the slot into the neighbouring ranks' frames."""

    def __init__(self, peer_ptrs, offset, mc_ptr=0, capacity=None):
        self.sink = (list(peer_ptrs), int(offset), int(mc_ptr or 0), None if capacity is None else int(capacity))

    def __enter__(self):
        self.prev = getattr(_tls, 'frame_sink', None)
        _tls.frame_sink = self.sink
        return self

    def __exit__(self, *exc):
        _tls.frame_sink = self.prev
        return False


def _current_frame_sink():
    return getattr(_tls, 'frame_sink', None)


def torgb_finish(raw, bias, clamp, img_prev, out_nchw=False, group=None):
    """raw [B,H,W,C]; img_prev [B,H/2,W/2,C] NHWC or None -> img [B,H,W,C] NHWC (or [B,C,H,W] planar)."""
    st = _enter(raw)
    B, H, W, Cc = raw.shape
    if out_nchw:
        out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=raw.device)
    else:
        out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=raw.device)
    if img_prev is not None:
        assert img_prev.is_contiguous() and tuple(img_prev.shape) == (B, H // 2, W // 2, Cc), (img_prev.shape, raw.shape)
    p = _C.TorgbParams(_p(raw), raw.stride(2), _p(bias), float(-1 if clamp is None else clamp), _p(img_prev), _p(out),
                       B, H, W, Cc, 1 if out_nchw else 0)
    if group is not None:
        p.groups, p.imgs_per_group = int(group[0]), int(group[1])
    sink = _current_frame_sink() if out_nchw else None
    if sink is not None:
        ptrs, off, mc, cap = sink
        if cap is not None and out.numel() != cap:
            raise RuntimeError(f'frame_sink: this call produces {out.numel()} image elements but the gathered buffer has slots of {cap} '
                               '(batch or resolution differs from the PeerFrameGather it was built for)')
        if mc:
            p.mc_out = mc
        else:
            for k, q in enumerate(ptrs):
                p.peer_out[k] = q
            p.n_peers = len(ptrs)
        p.peer_offset = off
    _C.check(_C.lib().ia_torgb_finish(C.byref(p), st), 'ia_torgb_finish')
    return out


# ---------------------------------------------------------------------------------------------------
# rasterize / stitch pieces
# ---------------------------------------------------------------------------------------------------
def fill_mouth(uv_or_alpha, upper_row0=87):
    """alpha given either as uvcoords_image [B,H,W,3] (mask in channel 2) or as [B,1,H,W]; returns
    (full_alpha, mouth, upper_alpha) as contiguous [B,H,W] tensors."""
    st = _enter(uv_or_alpha)
    t = uv_or_alpha if uv_or_alpha.dtype == torch.float32 else uv_or_alpha.float()
    if t.ndim == 4 and t.shape[-1] == 3 and t.shape[1] != 1:
        t = t.contiguous()
        B, H, W, _ = t.shape
        base, a_stride, a_batch = t.data_ptr() + 2 * 4, 3, H * W * 3
    else:
        t = t.contiguous()
        B, _, H, W = t.shape
        base, a_stride, a_batch = t.data_ptr(), 1, H * W
    full = torch.empty((B, H, W), dtype=torch.float32, device=t.device)
    mouth = torch.empty_like(full)
    upper = torch.empty_like(full)
    _C.check(_C.lib().ia_fill_mouth(base, a_stride, a_batch, B, H, W, int(upper_row0), _p(full), _p(mouth), _p(upper), st),
             'ia_fill_mouth')
    return full, mouth, upper


def grid_sample_nhwc(x_nhwc, grid, g_ld=None):
    """x [B,Hi,Wi,C] contiguous; grid [B,Ho,Wo,>=2] (last-dim stride 1) -> [B,Ho,Wo,C]."""
    st = _enter(x_nhwc)
    B, Hi, Wi, Cc = x_nhwc.shape
    assert x_nhwc.is_contiguous()
    grid = grid if grid.dtype == torch.float32 else grid.float()
    if not grid.is_contiguous():
        grid = grid.contiguous()
    _, Ho, Wo, gl = grid.shape
    out = torch.empty((B, Ho, Wo, Cc), dtype=torch.float32, device=x_nhwc.device)
    _C.check(_C.lib().ia_grid_sample(_p(x_nhwc), B, Hi, Wi, Cc, Cc, _p(grid), gl, Ho, Wo, _p(out), Cc, st), 'ia_grid_sample')
    return out


_AA_CACHE = {}


def aa_tables(in_size, out_size, device):
    """Tap tables of ATen's _upsample_bilinear2d_aa for one axis (float32 arithmetic as in
    aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa)."""
    key = (in_size, out_size, str(device))
    hit = _AA_CACHE.get(key)
    if hit is not None:
        return hit
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = f32(1.0) * scale if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    max_taps = int(math.ceil(float(support))) * 2 + 1
    starts = np.zeros(out_size, dtype=np.int32)
    counts = np.zeros(out_size, dtype=np.int32)
    weights = np.zeros((out_size, max_taps), dtype=np.float32)
    for i in range(out_size):
        center = scale * f32(i + 0.5)
        xmin = max(int(center - support + f32(0.5)), 0)
        xsize = min(int(center + support + f32(0.5)), in_size) - xmin
        xsize = max(min(xsize, max_taps), 0)
        ws = np.zeros(max_taps, dtype=np.float32)
        total = f32(0.0)
        for j in range(xsize):
            x = (f32(j + xmin) - center + f32(0.5)) * invscale
            x = -x if x < 0 else x
            w = f32(1.0) - x if x < 1.0 else f32(0.0)
            ws[j] = w
            total = f32(total + w)
        if total != 0:
            ws[:xsize] = ws[:xsize] / total
        starts[i], counts[i] = xmin, xsize
        weights[i] = ws
    out = (torch.from_numpy(starts).to(device), torch.from_numpy(counts).to(device),
           torch.from_numpy(weights).to(device), max_taps)
    _AA_CACHE[key] = out
    return out


def resize_aa(x_nhwc, oh, ow, crop=None, out=None, out_origin=(0, 0)):
    """Antialiased bilinear resize of (a crop of) x [B,H,W,C] to [oh,ow], optionally pasted into ``out`` at out_origin."""
    st = _enter(x_nhwc)
    B, H, W, Cc = x_nhwc.shape
    assert x_nhwc.stride(3) == 1 and x_nhwc.stride(1) == W * x_nhwc.stride(2) and x_nhwc.stride(0) == H * x_nhwc.stride(1)
    y0, y1, x0, x1 = (0, H, 0, W) if crop is None else crop
    ih, iw = y1 - y0, x1 - x0
    ys, yc, yw, ymt = aa_tables(ih, oh, x_nhwc.device)
    xs, xc, xw, xmt = aa_tables(iw, ow, x_nhwc.device)
    if out is None:
        out = torch.empty((B, oh, ow, Cc), dtype=torch.float32, device=x_nhwc.device)
    OB, OHh, OWw, OC = out.shape
    assert OB == B and out.stride(3) == 1 and out.stride(1) == OWw * out.stride(2) and out.stride(0) == OHh * out.stride(1)
    p = _C.ResizeParams(_p(x_nhwc), x_nhwc.stride(2), H, W, y0, x0, _p(out), out.stride(2), OHh, OWw, out_origin[0], out_origin[1],
                        B, Cc, oh, ow, _p(ys), _p(yc), _p(yw), ymt, _p(xs), _p(xc), _p(xw), xmt)
    _C.check(_C.lib().ia_resize_aa(C.byref(p), st), 'ia_resize_aa')
    return out


def lerp_alpha(a, b, alpha, out=None):
    """out = a*alpha + b*(1-alpha); a,b,out [B,H,W,C] NHWC views (unit channel stride), alpha [B,H,W] view."""
    st = _enter(a)
    B, H, W, Cc = a.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=a.device)
    for t in (a, b, out):
        assert t.stride(3) == 1 and tuple(t.shape) == (B, H, W, Cc)
    assert tuple(alpha.shape) == (B, H, W)
    p = _C.LerpParams(_p(a), a.stride(2), a.stride(1), a.stride(0), _p(b), b.stride(2), b.stride(1), b.stride(0),
                      _p(alpha), alpha.stride(2), alpha.stride(1), alpha.stride(0),
                      _p(out), out.stride(2), out.stride(1), out.stride(0), B, H, W, Cc)
    _C.check(_C.lib().ia_lerp_alpha(C.byref(p), st), 'ia_lerp_alpha')
    return out


def raster_level(tex_nhwc, uv, stat_nhwc, crop, alpha_r, res):
    """One rasterize pyramid level (triplane_v20.py:328-338) fused: aa_resize(grid_sample(tex, uv))*alpha +
    aa_resize(static[crop])*(1-alpha) -> [B,res,res,C] without the 256^2 x C intermediate.
    tex [B,Ht,Wt,C] contiguous; uv [B,UH,UW,>=2] contiguous; stat [B,SH,SW,>=C] NHWC view; crop = (y0,y1,x0,x1); alpha_r [B,res,res]."""
    st = _enter(tex_nhwc)
    B, Ht, Wt, Cc = tex_nhwc.shape
    assert tex_nhwc.is_contiguous() and uv.is_contiguous() and alpha_r.is_contiguous() and Cc % 4 == 0
    _, UH, UW, ul = uv.shape
    SB, SH, SW, SC = stat_nhwc.shape
    assert SC >= Cc and stat_nhwc.stride(3) == 1 and stat_nhwc.stride(1) == SW * stat_nhwc.stride(2) and stat_nhwc.stride(0) == SH * stat_nhwc.stride(1)
    y0, y1, x0, x1 = crop
    dev = tex_nhwc.device
    uxs, uxc, uxw, uxm = aa_tables(UW, res, dev)
    uys, uyc, uyw, uym = aa_tables(UH, res, dev)
    sxs, sxc, sxw, sxm = aa_tables(x1 - x0, res, dev)
    sys_, syc, syw, sym = aa_tables(y1 - y0, res, dev)
    tmp = torch.empty((B, UH, res, Cc), dtype=torch.float32, device=dev)
    out = torch.empty((B, res, res, Cc), dtype=torch.float32, device=dev)
    p = _C.RasterLevelParams(_p(tex_nhwc), Ht, Wt, Cc, _p(uv), ul, UH, UW, _p(tmp), _p(stat_nhwc), stat_nhwc.stride(2), SH, SW, y0, x0,
                             _p(alpha_r), _p(out), Cc, B, res,
                             _p(uxs), _p(uxc), _p(uxw), uxm, _p(uys), _p(uyc), _p(uyw), uym,
                             _p(sxs), _p(sxc), _p(sxw), sxm, _p(sys_), _p(syc), _p(syw), sym)
    _C.check(_C.lib().ia_raster_level(C.byref(p), st), 'ia_raster_level')
    return out


# ---------------------------------------------------------------------------------------------------
# renderer
# ---------------------------------------------------------------------------------------------------
def ray_march(colors, densities, depths, white_back=False):
    """MipRayMarcher2.run_forward (ray_marcher.py:25-57): colors [B,R,S,C], densities / depths [B,R,S,1] sorted along S ->
    (rgb [B,R,C], depth [B,R,1] clamped to the global depth range, weights [B,R,S-1,1])."""
    st = _enter(colors)
    colors, densities, depths = _f32c(colors), _f32c(densities), _f32c(depths)
    B, R, S, Cc = colors.shape
    assert densities.numel() == B * R * S and depths.numel() == B * R * S
    dev = colors.device
    rgb = torch.empty((B, R, Cc), dtype=torch.float32, device=dev)
    depth = torch.empty((B, R, 1), dtype=torch.float32, device=dev)
    weights = torch.empty((B, R, S - 1, 1), dtype=torch.float32, device=dev)
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    _C.check(_C.lib().ia_ray_march(_p(colors), _p(densities), _p(depths), B * R, S, Cc, 1 if white_back else 0, _p(rgb), _p(depth), _p(weights),
                                   _p(mm), st), 'ia_ray_march')
    _C.check(_C.lib().ia_depth_clamp(_p(depth), depth.numel(), _p(mm), st), 'ia_depth_clamp')
    return rgb, depth, weights


def stitch_planes(plane_img, stitch, alpha, origin, fp16=False):
    """Plane stitch (triplane_v20.py:119-128) in one pass: a copy of plane_img [B,H,W,C] fp32 NHWC (view with pixel stride >= C)
    whose plane-0 channels [0,32) inside the window at ``origin`` = (y0, x0) are stitch*alpha + plane*(1-alpha); stitch
    [B,h,w,32], alpha [B,h,w].  fp16=True writes the renderer's half-precision storage format."""
    st = _enter(plane_img)
    B, H, W, Cc = plane_img.shape
    assert plane_img.stride(3) == 1 and plane_img.stride(1) == W * plane_img.stride(2) and plane_img.stride(0) == H * plane_img.stride(1)
    stitch, alpha = _f32c(stitch), _f32c(alpha)
    _, wh, ww, sc = stitch.shape
    assert sc == 32 and tuple(alpha.shape) == (B, wh, ww)
    out = torch.empty((B, H, W, Cc), dtype=torch.float16 if fp16 else torch.float32, device=plane_img.device)
    p = _C.StitchParams(_p(plane_img), plane_img.stride(2), B, H, W, Cc, _p(stitch), _p(alpha), int(origin[0]), int(origin[1]), wh, ww,
                        _p(out), FMT_F16X1 if fp16 else 0)
    _C.check(_C.lib().ia_stitch_planes(C.byref(p), st), 'ia_stitch_planes')
    return out


def ray_sampler(cam, res):
    st = _enter(cam)
    cam = _f32c(cam)
    B = cam.shape[0]
    o = torch.empty((B, res * res, 3), dtype=torch.float32, device=cam.device)
    d = torch.empty_like(o)
    _C.check(_C.lib().ia_ray_sampler(_p(cam), cam.stride(0), B, res, _p(o), _p(d), st), 'ia_ray_sampler')
    return o, d


def render(planes_nhwc, cam, res, Dc, Df, jitter, u, box_warp, white_back, w1, b1, w2, b2, rays=None, mlp_fmt=0):
    """planes [B,PH,PW,>=96] NHWC fp32 (plane p = channels 32p..32p+31); cam [B,>=25] (or rays=(origins, dirs)
    [B,res*res,3] with cam=None); returns feat [B,res,res,32], depth [B,res,res] (clamped), wsum [B,res,res]."""
    st = _enter(planes_nhwc)
    B, PH, PW, PC = planes_nhwc.shape
    assert planes_nhwc.is_contiguous() and PC >= 96 and planes_nhwc.dtype in (torch.float32, torch.float16)
    planes_fmt = FMT_F16X1 if planes_nhwc.dtype == torch.float16 else 0
    rays_o = rays_d = None
    if rays is not None:
        rays_o, rays_d = _f32c(rays[0]), _f32c(rays[1])
        assert rays_o.numel() == B * res * res * 3 and rays_d.numel() == rays_o.numel()
    else:
        cam = _f32c(cam)
    jitter = _f32c(jitter)
    assert jitter.numel() == B * res * res * Dc, (jitter.shape, B, res, Dc)
    if u is not None:
        u = _f32c(u)
        assert u.numel() == B * res * res * Df
    dev = planes_nhwc.device
    near_far = torch.empty(4, dtype=torch.float32, device=dev)
    if rays is not None:
        _C.check(_C.lib().ia_ray_bounds_from_origins(_p(rays_o), rays_o.numel() // 3, _p(near_far), st), 'ia_ray_bounds_from_origins')
    else:
        _C.check(_C.lib().ia_ray_bounds(_p(cam), cam.stride(0), B, _p(near_far), st), 'ia_ray_bounds')
    feat = torch.empty((B, res, res, 32), dtype=torch.float32, device=dev)
    depth = torch.empty((B, res, res), dtype=torch.float32, device=dev)
    wsum = torch.empty_like(depth)
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    # per-call scratch for the staged decoder fragments (allocated on the launching stream by torch's caching allocator:
    # the library keeps no mutable device state, so renders on different streams cannot race)
    scratch = torch.empty(int(_C.lib().ia_render_scratch_bytes()), dtype=torch.uint8, device=dev)
    w1, b1, w2, b2 = _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2)
    p = _C.RenderParams(_p(planes_nhwc), PC, B, PH, PW, _p(cam), (cam.stride(0) if cam is not None else 0), _p(rays_o), _p(rays_d), res, Dc, Df, _p(jitter), _p(u),
                        float(box_warp), 1 if white_back else 0, _p(near_far), _p(w1), _p(b1), _p(w2), _p(b2),
                        _p(feat), _p(depth), _p(wsum), _p(mm), _p(scratch), int(mlp_fmt), planes_fmt)
    _C.check(_C.lib().ia_render(C.byref(p), st), 'ia_render')
    _C.check(_C.lib().ia_depth_clamp(_p(depth), depth.numel(), _p(mm), st), 'ia_depth_clamp')
    return feat, depth, wsum


# ---------------------------------------------------------------------------------------------------
# inversion-encoder pieces (csrc/ia_encoder.cu): everything reads through ia_view, so torch views (permute, slicing,
# expand) describe NCHW inputs, channel slices, stride-2 subsampling and batch broadcast without copies
# ---------------------------------------------------------------------------------------------------
def make_view(t, ps=1):
    """t: fp32 CUDA tensor indexed [B,H,W,C] (any strides, e.g. an NCHW tensor permuted) -> (_C.View, (B, H*ps, W*ps, C/ps^2)).
    ps > 1 reads it through torch.nn.PixelShuffle(ps)."""
    _require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    B, H, W, Cc = t.shape
    assert Cc % (ps * ps) == 0
    v = _C.View(t.data_ptr(), Cc // (ps * ps), ps, t.stride(3), t.stride(2), t.stride(1), t.stride(0))
    v._keep = t   # keep a converted temporary alive until the launch is enqueued
    return v, (B, H * ps, W * ps, Cc // (ps * ps))


def _as_view(src):
    """src: tensor [B,H,W,C] or (tensor, ps)."""
    return make_view(src[0], src[1]) if isinstance(src, (tuple, list)) else make_view(src, 1)


def enc_chan_stats(src):
    v, (B, H, W, Cc) = _as_view(src)
    st = _enter(v._keep)
    sums = torch.empty(2 * Cc, dtype=torch.float64, device=v._keep.device)
    _C.check(_C.lib().ia_enc_chan_stats(C.byref(v), B, H, W, _p(sums), st), 'ia_enc_chan_stats')
    return sums, B * H * W


_BN_SCRATCH = {}


class _FoldCache(_EngineCache):
    """Folded eval-mode BatchNorm (scale, shift) hung on the module; dropped on deepcopy / pickle like every engine cache."""

    def __init__(self, key, scale, shift):
        self.key, self.scale, self.shift = key, scale, shift


def enc_bn_fold(bn, srcs):
    """torch.nn.BatchNorm2d ``bn`` applied to cat(srcs, channel) -> per-channel (scale, shift) fp32 [C].
    Train mode (or no running statistics): batch statistics of the sources, running statistics updated as torch does."""
    dev = bn.weight.device if bn.weight is not None else srcs[0].device
    Cn = bn.num_features
    training = bn.training or not bn.track_running_stats
    if not training:
        # eval mode: (scale, shift) depend on the parameters and running statistics only -- folded once, reused until one of them
        # changes (storage or version)
        key = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None) + (float(bn.eps), str(dev))
        c = bn.__dict__.get('_ia_fold')
        if c is not None and c.key == key:
            return c.scale, c.shift
    scale = torch.empty(Cn, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    sums, count = None, 0
    if training and len(srcs) == 1 and enc_epilogue_fusion() and Cn <= 2048:
        # one launch: channel sums + the fold by the last CTA to arrive (persistent zeroed scratch per device / stream / capture scope)
        v, (B, H, W, Cc) = _as_view(srcs[0])
        assert Cc == Cn, (Cc, Cn)
        st = _enter(v._keep)
        key = _scratch_key(dev, st)
        buf = _BN_SCRATCH.get(key)
        if buf is None:
            buf = (torch.zeros(4096, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
            _BN_SCRATCH[key] = buf
        track = bn.track_running_stats and bn.running_mean is not None
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        if track and bn.momentum is None:
            momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
        _C.check(_C.lib().ia_enc_bn_stats_fold(C.byref(v), B, H, W, _p(buf[0]), _p(buf[1]), _p(bn.weight), _p(bn.bias),
                                               _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None, momentum,
                                               float(bn.eps), _p(scale), _p(shift), st), 'ia_enc_bn_stats_fold')
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        return scale, shift
    if training:
        if len(srcs) == 1:
            sums, count = enc_chan_stats(srcs[0])
        else:
            sums = torch.empty(2, Cn, dtype=torch.float64, device=dev)
            c0 = 0
            for s in srcs:
                part, count = enc_chan_stats(s)
                cs = part.numel() // 2
                sums[:, c0:c0 + cs] = part.view(2, cs)
                c0 += cs
            assert c0 == Cn
        assert sums.numel() == 2 * Cn, (sums.numel(), Cn)
    st = _enter(scale)
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = 0.1 if bn.momentum is None else float(bn.momentum)
    if training and track and bn.momentum is None:   # cumulative moving average
        momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
    _C.check(_C.lib().ia_enc_bn_fold(_p(sums), int(count), _p(bn.weight), _p(bn.bias), _p(bn.running_mean) if track else None,
                                     _p(bn.running_var) if track else None, 1 if training else 0, momentum, float(bn.eps), Cn,
                                     _p(scale), _p(shift), st), 'ia_enc_bn_fold')
    if training and track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    if not training:
        bn.__dict__['_ia_fold'] = _FoldCache(key, scale, shift)
    return scale, shift


def enc_prep(srcs, scale=None, shift=None, slope=None, lrelu=1.0, C_pad=None, want_split=True, want32=False):
    """cat(srcs) -> affine -> PReLU/leaky -> (Split [B,H,W,C_pad] bf16 hi/lo or None, fp32 NHWC [B,H,W,Ctot] or None)."""
    views = [_as_view(s) for s in srcs]
    B, H, W, _ = views[0][1]
    for _, shp in views:
        assert shp[:3] == (B, H, W), ('enc_prep: sources disagree', [s for _, s in views])
    Ctot = sum(shp[3] for _, shp in views)
    dev = views[0][0]._keep.device
    st = _enter(views[0][0]._keep)
    p = _C.EncPrepParams()
    for i, (v, _) in enumerate(views):
        p.src[i] = v
    p.nsrc = len(views)
    p.scale, p.shift, p.slope, p.lrelu = _p(scale), _p(shift), _p(slope), float(lrelu)
    sp = out32 = None
    if want_split:
        C_pad = _pad_to(Ctot, 64) if C_pad is None else C_pad
        sp = Split(torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=dev), torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=dev))
        p.hi, p.lo, p.C_pad = _p(sp.hi), _p(sp.lo), C_pad
    if want32:
        out32 = torch.empty((B, H, W, Ctot), dtype=torch.float32, device=dev)
        p.out32 = _p(out32)
    p.B, p.H, p.W = B, H, W
    _C.check(_C.lib().ia_enc_prep(C.byref(p), st), 'ia_enc_prep')
    return sp, out32


def enc_conv(a, conv, alg_stride=None):
    """a: Split; conv: torch.nn.Conv2d holder (3x3 pad 1 or 1x1, stride applied by the caller as a [::s, ::s] view of the
    result) -> raw fp32 accumulators [B,H,W,Cout] (bias not added).  alg_stride: stride used by the algorithmic-FLOP counter
    (default conv.stride; 1 when the caller already sub-sampled the operand)."""
    pack = ConvPack.current(conv, '_ia_pack', conv.weight, need_wsq=False)
    assert (pack.kh, pack.kw) in ((1, 1), (3, 3)) and a.C_pad == pack.Cin_pad, (pack.kh, a.C_pad, pack.Cin_pad)
    B, H, W, _ = a.hi.shape
    raw = torch.empty((B, H, W, pack.Cout), dtype=torch.float32, device=a.hi.device)
    stride = alg_stride if alg_stride is not None else (conv.stride[0] if isinstance(conv.stride, (tuple, list)) else int(conv.stride))
    conv_same(a.hi, a.lo, pack, pack.Cin_pad, raw, mode=0, alg_stride=stride)
    return raw


def enc_conv_act(a, conv, C_pad, slope=None, lrelu=None, want32=False):
    """conv (3x3 pad 1 or 1x1, stride 1) + bias + PReLU(slope) / LeakyReLU(lrelu) with the activation applied in the convolution's
    epilogue, which emits the next convolution's operand directly: -> Split [B,H,W,C_pad] (+ fp32 [B,H,W,Cout] when want32)."""
    pack = ConvPack.current(conv, '_ia_pack', conv.weight, need_wsq=False)
    assert (pack.kh, pack.kw) in ((1, 1), (3, 3)) and a.C_pad == pack.Cin_pad and pack.Cout % 4 == 0
    B, H, W, _ = a.hi.shape
    sp = new_split(B, H, W, C_pad, a.hi.device, C=pack.Cout)
    out32 = torch.empty((B, H, W, pack.Cout), dtype=torch.float32, device=a.hi.device) if want32 else None
    if slope is not None:
        conv_same(a.hi, a.lo, pack, pack.Cin_pad, out32, bias=conv.bias, act='prelu', slope=slope, gain=1.0, mode=1, e1=(sp, None))
    else:
        conv_same(a.hi, a.lo, pack, pack.Cin_pad, out32, bias=conv.bias, act='lrelu', alpha=float(lrelu), gain=1.0, mode=1, e1=(sp, None))
    return (sp, out32) if want32 else sp


def enc_epilogue_fusion():
    """Encoder convolutions apply the PReLU / LeakyReLU that follows them in their own epilogue and emit the next operand
    (IA_ENC_FUSE=0: separate ia_enc_prep passes, the round-1 formulation)."""
    return os.environ.get('IA_ENC_FUSE', '1') != '0' and _conv_impl == 'tc'


def enc_affine_act(x, scale=None, shift=None, slope1=None, slope2=None, act='linear', alpha=0.0, gate=None, res=None,
                   res_scale=None, res_shift=None, out=None, emit=None):
    """y = gate * act2(act1(x*scale + shift)) + (res*res_scale + res_shift); x/res: tensors [B,H,W,C] (any strides) or
    (tensor, ps).  out: optional [B,H,W,C] destination with unit channel stride and dense rows (a column slice of a
    wider NHWC buffer is fine)."""
    xv, (B, H, W, Cc) = _as_view(x)
    st = _enter(xv._keep)
    p = _C.EncAffineParams()
    p.x = xv
    p.scale, p.shift, p.slope1, p.slope2 = _p(scale), _p(shift), _p(slope1), _p(slope2)
    p.act, p.alpha, p.gate = ACT_IDS[act], float(alpha), _p(gate)
    if res is not None:
        rv, rshape = _as_view(res)
        if rshape[0] == 1 and B > 1:
            rv.s_img = 0
        assert rshape[1:] == (H, W, Cc), (rshape, (B, H, W, Cc))
        p.res = rv
    p.res_scale, p.res_shift = _p(res_scale), _p(res_shift)
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=xv._keep.device)
    assert tuple(out.shape) == (B, H, W, Cc) and (Cc == 1 or out.stride(3) == 1), (out.shape, out.stride())
    # the destination must be pixel-linear: address(b, y, x) = ((b*H + y)*W + x) * y_ld  (size-1 dims carry no constraint)
    y_ld = out.stride(2) if W > 1 else (out.stride(1) if H > 1 else (out.stride(0) if B > 1 else Cc))
    assert (W == 1 or H == 1 or out.stride(1) == W * y_ld) and (B == 1 or H * W == 1 or out.stride(0) == H * W * y_ld), (out.shape, out.stride())
    p.y, p.y_ld = _p(out), y_ld
    p.B, p.H, p.W, p.C = B, H, W, Cc
    sp = None
    if emit is not None:      # (e_scale, e_shift, C_pad): also write split(y * e_scale + e_shift), the next convolution's operand
        e_scale, e_shift, C_pad = emit
        sp = new_split(B, H, W, C_pad, out.device)
        p.e_scale, p.e_shift, p.e_hi, p.e_lo, p.e_C_pad = _p(e_scale), _p(e_shift), _p(sp.hi), _p(sp.lo), C_pad
    _C.check(_C.lib().ia_enc_affine_act(C.byref(p), st), 'ia_enc_affine_act')
    return out if emit is None else (out, sp)


def enc_global_pool(x, scale=None, shift=None):
    xv, (B, H, W, Cc) = _as_view(x)
    st = _enter(xv._keep)
    pooled = torch.empty((B, Cc), dtype=torch.float32, device=xv._keep.device)
    _C.check(_C.lib().ia_enc_global_pool(C.byref(xv), _p(scale), _p(shift), B, H, W, _p(pooled), st), 'ia_enc_global_pool')
    return pooled


_SE_SCRATCH = {}


def enc_se_gate(x, scale, shift, w1, w2):
    """SEModule gate in one launch: x [B,H,W,C] view, (scale, shift) the folded BatchNorm in front of it (or None), w1 [Cr,C], w2 [C,Cr]
    -> gate [B,C].  Zero-initialised scratch per (device, stream), left zeroed by every launch."""
    xv, (B, H, W, Cc) = _as_view(x)
    st = _enter(xv._keep)
    dev = xv._keep.device
    key = _scratch_key(dev, st)
    buf = _SE_SCRATCH.get(key)
    if buf is None or buf[0].numel() < B * Cc or buf[1].numel() < B:
        buf = (torch.zeros(max(B * Cc, 64 * 1024), dtype=torch.float32, device=dev), torch.zeros(max(B, 256), dtype=torch.int32, device=dev))
        _SE_SCRATCH[key] = buf
    gate = torch.empty((B, Cc), dtype=torch.float32, device=dev)
    Cr = w1.shape[0]
    assert tuple(w1.shape) == (Cr, Cc) and tuple(w2.shape) == (Cc, Cr) and w1.is_contiguous() and w2.is_contiguous()
    _C.check(_C.lib().ia_enc_se_gate(C.byref(xv), _p(scale), _p(shift), B, H, W, _p(w1), _p(w2), Cr, _p(buf[0]), _p(buf[1]), _p(gate), st),
             'ia_enc_se_gate')
    return gate


def enc_avgpool(x, k):
    xv, (B, H, W, Cc) = _as_view(x)
    st = _enter(xv._keep)
    y = torch.empty((B, H // k, W // k, Cc), dtype=torch.float32, device=xv._keep.device)
    _C.check(_C.lib().ia_enc_avgpool(C.byref(xv), B, H, W, int(k), _p(y), st), 'ia_enc_avgpool')
    return y


def enc_upsample_add(x, lateral):
    """bilinear(align_corners=True) upsample of x [B,h,w,C] to lateral's size, plus lateral (both contiguous NHWC)."""
    st = _enter(x)
    x, lateral = _f32c(x), _f32c(lateral)
    B, h, w, Cc = x.shape
    _, H, W, _ = lateral.shape
    y = torch.empty_like(lateral)
    _C.check(_C.lib().ia_enc_upsample_add(_p(x), B, h, w, Cc, _p(lateral), H, W, _p(y), st), 'ia_enc_upsample_add')
    return y


def enc_gru_gate0(raw, bias, h):
    """raw [B,H,W,2C] (ih accumulators) -> (r*h, z) each [B,H,W,C]."""
    st = _enter(raw)
    B, H, W, C2 = raw.shape
    Cc = C2 // 2
    rh = torch.empty((B, H, W, Cc), dtype=torch.float32, device=raw.device)
    z = torch.empty_like(rh)
    assert raw.is_contiguous() and (h is None or h.is_contiguous())
    _C.check(_C.lib().ia_enc_gru_gate(0, _p(raw), _p(bias), _p(h), _p(rh), _p(z), None, B * H * W, Cc, st), 'ia_enc_gru_gate')
    return rh, z


def enc_gru_gate1(raw, bias, h, z):
    st = _enter(raw)
    B, H, W, Cc = raw.shape
    h_out = torch.empty_like(raw)
    assert raw.is_contiguous() and z.is_contiguous() and (h is None or h.is_contiguous())
    _C.check(_C.lib().ia_enc_gru_gate(1, _p(raw), _p(bias), _p(h), None, _p(z), _p(h_out), B * H * W, Cc, st), 'ia_enc_gru_gate')
    return h_out


def sft_half(x_nhwc, scale, shift):
    """In place CS-SFT on the second half of the channels of x [B,H,W,C] (networks_stylegan2_new.py:448-452);
    scale/shift: [1|B,H,W,C/2]-indexed tensors (any strides)."""
    st = _enter(x_nhwc)
    B, H, W, Cc = x_nhwc.shape
    assert x_nhwc.stride(3) == 1 and x_nhwc.stride(1) == W * x_nhwc.stride(2) and x_nhwc.stride(0) == H * x_nhwc.stride(1)
    sv, sshape = make_view(scale)
    hv, hshape = make_view(shift)
    assert sshape[1:] == (H, W, Cc // 2) and hshape[1:] == (H, W, Cc // 2), (sshape, hshape, x_nhwc.shape)
    if sshape[0] == 1:
        sv.s_img = 0
    if hshape[0] == 1:
        hv.s_img = 0
    _C.check(_C.lib().ia_sft_half(_p(x_nhwc), x_nhwc.stride(2), C.byref(sv), C.byref(hv), B, H, W, Cc, st), 'ia_sft_half')
    return x_nhwc


# ---------------------------------------------------------------------------------------------------
# Mix-Transformer pieces of the improved one-shot encoder (csrc/ia_vit.cu).  Tokens are [B,H,W,C] fp32 maps.
# ---------------------------------------------------------------------------------------------------
def linear_pack(owner, weight, attr='_ia_pack'):
    """ConvPack of an nn.Linear weight [Out,In] seen as a 1x1 convolution (cached on ``owner``, repacked when the parameter changes)."""
    return ConvPack.current(owner, attr, weight.detach().view(weight.shape[0], weight.shape[1], 1, 1), need_wsq=False)


class _Im2colPack(_EngineCache):
    """ConvPack of a k x k nn.Conv2d weight laid out for the im2col GEMM: [O][kh][kw][I] flattened to a 1x1 convolution with
    kh*kw*I input channels (the K order ia_enc_im2col writes)."""

    def __init__(self, weight):
        O, I, kh, kw = weight.shape
        w2 = weight.detach().permute(0, 2, 3, 1).reshape(O, kh * kw * I, 1, 1).contiguous()
        self.pack = ConvPack(w2, need_wsq=False)
        self.k = kh
        self.key = (weight.data_ptr(), weight._version, str(weight.device))

    @staticmethod
    def current(conv):
        w = conv.weight
        assert w.shape[2] == w.shape[3] and conv.groups == 1
        c = conv.__dict__.get('_ia_pack_i2c')
        if c is None or c.key != (w.data_ptr(), w._version, str(w.device)):
            c = _Im2colPack(w)
            conv.__dict__['_ia_pack_i2c'] = c
        return c.pack


def im2col_pack(conv):
    return _Im2colPack.current(conv)


def enc_im2col(srcs, k, stride, pad, K_pad=None):
    """cat(srcs) [B,H,W,C] (tensors or (tensor, pixel_shuffle)) -> Split [B,OH,OW,K_pad] of k x k patches, K order (ky, kx, c)."""
    views = [_as_view(s) for s in srcs]
    B, H, W, _ = views[0][1]
    for _, shp in views:
        assert shp[:3] == (B, H, W), ('enc_im2col: sources disagree', [s for _, s in views])
    Ctot = sum(shp[3] for _, shp in views)
    dev = views[0][0]._keep.device
    st = _enter(views[0][0]._keep)
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    K_pad = _pad_to(k * k * Ctot, 64) if K_pad is None else K_pad
    sp = Split(torch.empty((B, OH, OW, K_pad), dtype=torch.bfloat16, device=dev), torch.empty((B, OH, OW, K_pad), dtype=torch.bfloat16, device=dev))
    p = _C.EncIm2colParams()
    for i, (v, _) in enumerate(views):
        p.src[i] = v
    p.nsrc = len(views)
    p.B, p.H, p.W, p.k, p.stride, p.pad, p.OH, p.OW = B, H, W, int(k), int(stride), int(pad), OH, OW
    p.hi, p.lo, p.K_pad = _p(sp.hi), _p(sp.lo), K_pad
    _C.check(_C.lib().ia_enc_im2col(C.byref(p), st), 'ia_enc_im2col')
    return sp


def enc_gemm(a, pack, emit_split=False):
    """a: Split [B,H,W,K_pad]; pack: ConvPack of a 1x1 convolution -> raw fp32 accumulators [B,H,W,Cout] (bias not added), or,
    with emit_split, the same values as the bf16 hi/lo operand [B,H,W,Cout] written by the GEMM's epilogue (no fp32 round trip)."""
    assert pack.taps == 1 and a.C_pad == pack.Cin_pad, (pack.taps, a.C_pad, pack.Cin_pad)
    B, H, W, _ = a.hi.shape
    if emit_split:
        assert pack.Cout % 64 == 0
        sp = new_split(B, H, W, pack.Cout, a.hi.device)
        conv_same(a.hi, a.lo, pack, pack.Cin_pad, None, mode=0, e1=(sp, None))
        return sp
    raw = torch.empty((B, H, W, pack.Cout), dtype=torch.float32, device=a.hi.device)
    conv_same(a.hi, a.lo, pack, pack.Cin_pad, raw, mode=0)
    return raw


def attention_tc_available(head_dim, q_bias, kv_bias):
    """Tensor-core attention (ia_attention_tc) covers head_dim 256 without q/kv bias -- transformer_block's configuration."""
    return head_dim == 256 and q_bias is None and kv_bias is None and os.environ.get('IA_ATTENTION', 'tc') != 'simt'


def attention_tc(q, kv, heads, scale, want32=False):
    """q: Split [B,H,W,C], kv: Split [B,h,w,2C] (the q / kv projections' emitted operands) -> Split [B,H,W,C] (+ fp32 copy)."""
    st = _enter(q.hi)
    B, H, W, Cc = q.hi.shape
    Bk, h, w, C2 = kv.hi.shape
    assert Bk == B and C2 == 2 * Cc and Cc == heads * 256 and q.fmt == FMT_BF16X3 and kv.fmt == FMT_BF16X3
    assert q.hi.is_contiguous() and kv.hi.is_contiguous()
    sp = new_split(B, H, W, Cc, q.hi.device)
    out32 = torch.empty((B, H, W, Cc), dtype=torch.float32, device=q.hi.device) if want32 else None
    p = _C.AttentionTcParams()
    p.q_hi, p.q_lo, p.q_ld, p.kv_hi, p.kv_lo, p.kv_ld = _p(q.hi), _p(q.lo), Cc, _p(kv.hi), _p(kv.lo), C2
    p.B, p.heads, p.head_dim, p.Nq, p.Nk, p.scale = B, int(heads), 256, H * W, h * w, float(scale)
    p.out32, p.out32_ld = _p(out32), Cc
    p.hi, p.lo, p.C_pad = _p(sp.hi), _p(sp.lo), Cc
    _C.check(_C.lib().ia_attention_tc(C.byref(p), st), 'ia_attention_tc')
    return (sp, out32) if want32 else sp


def layer_norm(x, ln, pre_bias=None, want_split=True, want32=False):
    """torch.nn.LayerNorm ``ln`` over the channels of x [B,H,W,C] fp32 (dense rows; + pre_bias[c] first) ->
    (Split [B,H,W,C_pad] or None, fp32 [B,H,W,C] or None)."""
    _require_cuda(x)
    st = _enter(x)
    B, H, W, Cc = x.shape
    assert x.dtype == torch.float32 and x.stride(3) == 1 and x.is_contiguous(), (x.dtype, x.stride())
    assert tuple(ln.normalized_shape) == (Cc,) and ln.elementwise_affine
    sp = out32 = None
    if want_split:
        C_pad = _pad_to(Cc, 64)
        sp = Split(torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=x.device), torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=x.device))
    if want32:
        out32 = torch.empty((B, H, W, Cc), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().ia_layer_norm(_p(x), Cc, _p(pre_bias), _p(ln.weight), _p(ln.bias), float(ln.eps), B * H * W, Cc, _p(out32), Cc,
                                    _p(sp.hi) if sp else None, _p(sp.lo) if sp else None, sp.C_pad if sp else 0, st), 'ia_layer_norm')
    return sp, out32


def attention(q, kv, heads, scale, q_bias=None, kv_bias=None, want32=False):
    """q [B,H,W,C] and kv [B,h,w,2C] fp32 raw projections (kv channels = (k | v), heads-major inside each half, the layout of
    mix_transformer.py:105) -> Split [B,H,W,C] of softmax(q k^T scale) v (the proj layer's operand) (+ fp32 copy)."""
    _require_cuda(q, kv)
    st = _enter(q)
    B, H, W, Cc = q.shape
    Bk, h, w, C2 = kv.shape
    assert Bk == B and C2 == 2 * Cc and Cc % heads == 0 and q.is_contiguous() and kv.is_contiguous()
    assert Cc % 64 == 0, 'attention: channel count must be a multiple of 64 (operand padding)'
    sp = Split(torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=q.device), torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=q.device))
    out32 = torch.empty((B, H, W, Cc), dtype=torch.float32, device=q.device) if want32 else None
    p = _C.AttentionParams()
    kvf = kv.view(-1)
    p.q, p.k, p.v, p.q_ld, p.k_ld, p.v_ld = _p(q), _p(kvf), _p(kvf[Cc:]), Cc, C2, C2
    if q_bias is not None:
        p.q_bias = _p(q_bias)
    if kv_bias is not None:
        kb = kv_bias.detach()
        p.k_bias, p.v_bias = _p(kb), _p(kb[Cc:])
    p.B, p.heads, p.head_dim, p.Nq, p.Nk, p.scale = B, int(heads), Cc // heads, H * W, h * w, float(scale)
    p.out32, p.out32_ld = _p(out32), Cc
    p.hi, p.lo, p.C_pad = _p(sp.hi), _p(sp.lo), Cc
    _C.check(_C.lib().ia_attention(C.byref(p), st), 'ia_attention')
    return (sp, out32) if want32 else sp


def dwconv_gelu(x, in_bias, dw):
    """gelu(depthwise3x3(x + in_bias) + bias): x [B,H,W,C] raw fc1 accumulators, dw: nn.Conv2d(C, C, 3, 1, 1, groups=C) -> Split."""
    _require_cuda(x)
    st = _enter(x)
    B, H, W, Cc = x.shape
    assert x.is_contiguous() and tuple(dw.weight.shape) == (Cc, 1, 3, 3)
    C_pad = _pad_to(Cc, 64)
    sp = Split(torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=x.device), torch.empty((B, H, W, C_pad), dtype=torch.bfloat16, device=x.device))
    _C.check(_C.lib().ia_dwconv_gelu(_p(x), _p(in_bias), _p(dw.weight), _p(dw.bias), B, H, W, Cc, None, _p(sp.hi), _p(sp.lo), C_pad, st),
             'ia_dwconv_gelu')
    return sp


# ---------------------------------------------------------------------------------------------------
# output stage
# ---------------------------------------------------------------------------------------------------
def layout_grid_u8(img, grid_w=None, grid_h=1):
    """[B,C,H,W] float image batch (any strides) -> uint8 [grid_h*H, grid_w*W, C] (reenact_avatar_next3d.py:117-131 with
    float_to_uint8=True, chw_to_hwc=True), one kernel; the D2H copy that follows moves 1 byte per value instead of 4."""
    _require_cuda(img)
    st = _enter(img)
    img = img if img.dtype == torch.float32 else img.float()
    B, Cc, H, W = img.shape
    if grid_w is None:
        grid_w = B // grid_h
    assert B == grid_w * grid_h
    out = torch.empty((grid_h * H, grid_w * W, Cc), dtype=torch.uint8, device=img.device)
    _C.check(_C.lib().ia_layout_grid_u8(_p(img), img.stride(0), img.stride(1), img.stride(2), img.stride(3), grid_h, grid_w, Cc, H, W,
                                        _p(out), st), 'ia_layout_grid_u8')
    return out


_SIDE_STREAMS = {}


def side_streams(device, n=2):
    """``n`` auxiliary CUDA streams per device (process-wide; kept off the modules so that they stay deep-copyable/picklable)."""
    key = str(torch.device(device))
    st = _SIDE_STREAMS.get(key, ())
    if len(st) < n:
        st = tuple(st) + tuple(torch.cuda.Stream(device=device) for _ in range(n - len(st)))
        _SIDE_STREAMS[key] = st
    return st[:n]


def _guard_public_entry_points():
    """Wrap every public function / method of this module that talks to the library in _device_guarded (see _enter)."""
    import inspect
    g = globals()
    for name, obj in list(g.items()):
        if inspect.isfunction(obj) and obj.__module__ == __name__ and not name.startswith('_') and '_enter' in obj.__code__.co_names:
            g[name] = _device_guarded(obj)
    ConvPack.__init__ = _device_guarded(ConvPack.__init__)
    StylePlan.run = _device_guarded(StylePlan.run)


_guard_public_entry_points()
