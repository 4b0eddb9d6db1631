"""GPU parity of the op-level C-ABI entry points (called through invertavatar_b200.runtime, the ctypes layer)
against the reference-minted golden vectors and the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import T, golden
from invertavatar_b200 import runtime as rt
from oracle import ops as o_ops
from oracle import stylegan2 as o_sg
from oracle import triplane as o_tp

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def close(a, b, atol, what=''):
    a = a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)
    b = b.detach().cpu().numpy() if hasattr(b, 'detach') else np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) if a.size else 0.0
    assert err <= atol, f'{what}: max-abs {err:.3e} > {atol:.1e}'
    return err


def test_bias_act_golden():
    g = golden('ops.npz')
    x, b = T(g['bias_act/x']).to(DEV), T(g['bias_act/b']).to(DEV)
    for act in rt.ACT_IDS:
        close(rt.bias_act(x, b, act=act), g[f'bias_act/{act}'], 2e-6, act)
    close(rt.bias_act(x, b, act='lrelu', alpha=0.1, gain=1.7, clamp=1.5), g['bias_act/lrelu_gain_clamp'], 1e-6)
    close(rt.bias_act(x, torch.arange(6.0, device=DEV), dim=3), g['bias_act/linear_dim3'], 1e-6)
    # channels-last input, no bias, empty tensor
    xc = x.contiguous(memory_format=torch.channels_last)
    close(rt.bias_act(xc, b, act='lrelu'), g['bias_act/lrelu'], 1e-6)
    close(rt.bias_act(x, None, act='relu'), o_ops.bias_act(x.cpu(), None, act='relu'), 1e-6)
    assert rt.bias_act(torch.empty(0, 5, 2, 2, device=DEV), b).numel() == 0
    with pytest.raises(RuntimeError):
        rt.bias_act(x.cpu(), b.cpu())          # no CPU fallback


def test_bias_act_large():
    x = torch.randn(2, 128, 64, 64, device=DEV)
    b = torch.randn(128, device=DEV)
    y = rt.bias_act(x, b, act='lrelu', clamp=256)
    close(y, o_ops.bias_act(x.cpu(), b.cpu(), act='lrelu', clamp=256), 1e-6)


def test_upfirdn2d_golden():
    g = golden('ops.npz')
    x, f = T(g['upfirdn2d/x']).to(DEV), T(g['upfirdn2d/f']).to(DEV)
    close(rt.upfirdn2d(x, f, up=(2, 2), padding=(2, 1, 2, 1), gain=4), g['upfirdn2d/up2'], 1e-6)
    close(rt.upfirdn2d(x, f, down=(2, 2), padding=(1, 1, 1, 1)), g['upfirdn2d/down2'], 1e-6)
    close(rt.upfirdn2d(x, f, padding=(2, 1, 2, 1)), g['upfirdn2d/filter'], 1e-6)
    close(rt.upfirdn2d(x, f, padding=(1, 1, 1, 1), gain=4), g['upfirdn2d/pad_fir'], 1e-6)
    close(rt.upfirdn2d(x, f, up=(3, 3), down=(2, 2), padding=(2, 1, 0, 3), flip_filter=True, gain=2), g['upfirdn2d/up3_down2_pad'], 1e-6)
    fa = T(g['upfirdn2d/fa']).to(DEV)
    close(rt.upfirdn2d(x, fa, up=(2, 1), down=(1, 2), padding=(1, 2, 2, 1)), g['upfirdn2d/asym'], 2e-6)
    close(rt.upfirdn2d(x, fa, up=(2, 1), down=(1, 2), padding=(1, 2, 2, 1), flip_filter=True), g['upfirdn2d/asym_flip'], 2e-6)
    close(rt.upfirdn2d(x, T(g['upfirdn2d/f_sep']).to(DEV), up=(2, 2), padding=(4, 3, 4, 3), gain=4), g['upfirdn2d/sep_up2'], 1e-6)
    close(rt.upfirdn2d(x, f, up=(2, 2), padding=(-1, 2, 3, -2)), g['upfirdn2d/negpad'], 1e-6)
    # channels-last strides are honoured (upfirdn2d.cpp:56-63)
    xc = x.contiguous(memory_format=torch.channels_last)
    y = rt.upfirdn2d(xc, f, up=(2, 2), padding=(2, 1, 2, 1), gain=4)
    close(y, g['upfirdn2d/up2'], 1e-6)


def test_mapping_pieces():
    x = torch.randn(5, 37, device=DEV)
    w = torch.randn(19, 37, device=DEV)
    b = torch.randn(19, device=DEV)
    y = rt.fully_connected(x, w, b, w_gain=0.01 / np.sqrt(37), b_gain=0.01, act='lrelu')
    close(y, o_sg.fully_connected(x.cpu(), w.cpu(), b.cpu(), activation='lrelu', lr_multiplier=0.01), 1e-6)
    x = torch.randn(11, 512, device=DEV)      # more than one batch chunk of 8
    w = torch.randn(512, 512, device=DEV)
    close(rt.fully_connected(x, w, None, w_gain=1 / np.sqrt(512)), o_sg.fully_connected(x.cpu(), w.cpu()), 2e-5)
    close(rt.normalize_2nd_moment(x), o_sg.normalize_2nd_moment(x.cpu()), 1e-5)
    wa = torch.randn(512, device=DEV)
    ws = rt.broadcast_truncate(x, wa, 14, psi=0.7, cutoff=9)
    ref = x.cpu().unsqueeze(1).repeat(1, 14, 1)
    ref[:, :9] = wa.cpu().lerp(ref[:, :9], 0.7)
    close(ws, ref, 1e-6)


def test_fill_mouth_golden():
    g = golden('stages.npz')
    a = T(g['fill_mouth/alpha']).to(DEV)
    full, mouth, upper = rt.fill_mouth(a)
    close(full.unsqueeze(1), g['fill_mouth/full'], 0)
    close(mouth.unsqueeze(1), g['fill_mouth/mouth'], 0)
    up = T(g['fill_mouth/mouth']).clone()
    up[:, :, :87] = 0
    close(upper.unsqueeze(1), (T(g['fill_mouth/alpha']) + up).clamp(0, 1), 0)
    # uvcoords_image layout (mask = channel 2 of [B,H,W,3])
    uv = torch.cat([torch.zeros(3, 256, 256, 2), T(g['fill_mouth/alpha'])[:, 0].unsqueeze(-1)], -1).to(DEV)
    full2, mouth2, _ = rt.fill_mouth(uv)
    close(full2, full, 0)
    close(mouth2, mouth, 0)


def test_fill_mouth_adversarial():
    """Spiral corridor: the flood has to wind through a long 1-pixel path (many row/column sweep iterations)."""
    a = torch.ones(1, 1, 64, 64)
    a[0, 0, 0, :] = 0
    y0, y1, x0, x1 = 2, 61, 2, 61
    a[0, 0, 0:3, 63] = 0
    # carve a serpentine path
    for r in range(2, 62, 4):
        a[0, 0, r, 1:63] = 0
        a[0, 0, r:r + 3, 62 if (r // 4) % 2 == 0 else 1] = 0
        if r + 2 < 64:
            a[0, 0, r + 2, 1:63] = 0
    a[0, 0, 1, 63] = 0
    a[0, 0, 40:44, 30:34] = 1
    a[0, 0, 41:43, 31:33] = 0.0    # enclosed 2x2 hole inside a solid 4x4 block sitting in the corridor area
    full, mouth = o_tp.fill_mouth(a.clone())
    f2, m2, _ = rt.fill_mouth(a.to(DEV))
    close(f2.unsqueeze(1), full, 0)
    close(m2.unsqueeze(1), mouth, 0)


def test_grid_sample_resize_lerp():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 12, 20, 24, generator=g)          # NCHW
    grid = torch.rand(2, 9, 11, 2, generator=g) * 2.4 - 1.2   # includes out-of-range samples
    ref = F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
    y = rt.grid_sample_nhwc(rt.to_nhwc(x.to(DEV)), grid.to(DEV))
    close(rt.from_nhwc(y), ref, 2e-6)
    for (ih, iw, oh, ow) in [(256, 256, 32, 32), (256, 256, 64, 64), (64, 64, 64, 64), (16, 16, 32, 32), (128, 128, 128, 128), (40, 24, 17, 31)]:
        x = torch.randn(2, 5, ih, iw, generator=g)
        ref = F.interpolate(x, size=(oh, ow), mode='bilinear', antialias=True)
        y = rt.resize_aa(rt.to_nhwc(x.to(DEV)), oh, ow)
        close(rt.from_nhwc(y), ref, 2e-6, f'resize {ih}x{iw}->{oh}x{ow}')
    # crop + paste window
    x = torch.randn(1, 8, 64, 64, generator=g)
    ref = F.interpolate(x[:, :, 14:46, 16:48], size=(64, 64), mode='bilinear', antialias=True)
    y = rt.resize_aa(rt.to_nhwc(x.to(DEV)), 64, 64, crop=(14, 46, 16, 48))
    close(rt.from_nhwc(y), ref, 2e-6)
    a, b = torch.randn(2, 6, 7, 5, generator=g), torch.randn(2, 6, 7, 5, generator=g)
    al = torch.rand(2, 6, 7, generator=g)
    y = rt.lerp_alpha(a.to(DEV), b.to(DEV), al.to(DEV))
    close(y, a * al.unsqueeze(-1) + b * (1 - al.unsqueeze(-1)), 1e-6)


def test_reference_python_api_golden():
    """The function-level mirror of torch_utils.ops (invertavatar_b200.ops, what the drop-in tree exports) against the
    reference-minted vectors: upsample2d / downsample2d / filter2d, filtered_lrelu, conv2d_resample, modulated_conv2d, fma,
    grid_sample."""
    from invertavatar_b200 import ops
    g = golden('ops.npz')
    x, f = T(g['upfirdn2d/x']).to(DEV), T(g['upfirdn2d/f']).to(DEV)
    close(ops.upsample2d(x, f), g['upfirdn2d/up2'], 1e-6)
    close(ops.downsample2d(x, f), g['upfirdn2d/down2'], 1e-6)
    close(ops.filter2d(x, f), g['upfirdn2d/filter'], 1e-6)
    close(ops.upsample2d(x, T(g['upfirdn2d/f_sep']).to(DEV)), g['upfirdn2d/sep_up2'], 1e-6)
    close(ops.upfirdn2d(x, T(g['upfirdn2d/fa']).to(DEV), up=[2, 1], down=[1, 2], padding=[1, 2, 2, 1]), g['upfirdn2d/asym'], 2e-6)
    xb, bb, ff = T(g['filtered_lrelu/x']).to(DEV), T(g['filtered_lrelu/b']).to(DEV), T(g['filtered_lrelu/f']).to(DEV)
    close(ops.filtered_lrelu(xb, fu=ff, fd=ff, b=bb, up=2, down=2, padding=3, clamp=0.9), g['filtered_lrelu/up2_down2'], 2e-6)
    close(ops.filtered_lrelu(xb, b=bb), g['filtered_lrelu/plain'], 1e-6)
    xm, w, s = T(g['modconv/x']).to(DEV), T(g['modconv/w']).to(DEV), T(g['modconv/s']).to(DEV)
    tol = 2e-4
    close(ops.conv2d_resample(xm, w, f=f, up=2, padding=1, flip_weight=False), g['conv2d_resample/up2'], tol * float(np.abs(g['conv2d_resample/up2']).max()))
    close(ops.conv2d_resample(xm, w, padding=1), g['conv2d_resample/same'], tol * float(np.abs(g['conv2d_resample/same']).max()))
    close(ops.conv2d(xm, w, padding=1), g['conv2d_resample/same'], tol * float(np.abs(g['conv2d_resample/same']).max()))
    close(ops.modulated_conv2d(xm, w, s, noise=T(g['modconv/noise']).to(DEV), padding=1), g['modconv/same'], tol * max(1.0, float(np.abs(g['modconv/same']).max())))
    close(ops.modulated_conv2d(xm, w, s, noise=T(g['modconv/noise2']).to(DEV), up=2, padding=1, resample_filter=f, flip_weight=False),
          g['modconv/up2'], tol * max(1.0, float(np.abs(g['modconv/up2']).max())))
    close(ops.modulated_conv2d(xm, T(g['modconv/w1']).to(DEV), s, demodulate=False), g['modconv/torgb'], tol * max(1.0, float(np.abs(g['modconv/torgb']).max())))
    a, b, c = torch.randn(3, 1, 5, device=DEV), torch.randn(1, 4, 5, device=DEV), torch.randn(3, 4, 1, device=DEV)
    close(ops.fma(a, b, c), (a * b + c), 1e-6)
    inp, grid = torch.randn(2, 5, 9, 7, device=DEV), torch.rand(2, 6, 4, 2, device=DEV) * 2.4 - 1.2
    close(ops.grid_sample(inp, grid), F.grid_sample(inp.cpu(), grid.cpu(), mode='bilinear', padding_mode='zeros', align_corners=False), 2e-6)
    with pytest.raises(NotImplementedError):
        ops.conv2d_resample(xm, w, f=f, down=2, padding=1)


@pytest.mark.parametrize('Cc,tr,res', [(32, 32, 32), (512, 32, 32), (512, 64, 64), (256, 128, 128), (128, 256, 256), (24, 16, 64)])
def test_raster_level_fused(Cc, tr, res):
    """One rasterize level (triplane_v20.py:328-338), fused two-pass kernel vs torch on CPU and vs the reference's rasterize golden
    path (grid_sample @256^2 -> antialiased resize -> alpha blend with the resized static crop)."""
    g = torch.Generator().manual_seed(Cc + res)
    B = 2
    tex = torch.randn(B, Cc, tr, tr, generator=g)
    uv = synth_uv(B)
    stat = torch.randn(B, Cc + 8, tr, tr, generator=g)        # wider than C: the kernel reads a channel slice
    bbox = [round(i * res / 256) for i in (57, 185, 64, 192)]
    sb = [round(i * tr / 256) for i in (57, 185, 64, 192)]
    alpha = F.interpolate(uv[..., 2].unsqueeze(1), size=(res, res), mode='bilinear', antialias=True)
    ri = F.grid_sample(tex, uv[..., :2], mode='bilinear', padding_mode='zeros', align_corners=False)
    rf = F.interpolate(ri, size=(res, res), mode='bilinear', antialias=True)
    sf = F.interpolate(stat[:, :Cc, sb[0]:sb[1], sb[2]:sb[3]], size=(res, res), mode='bilinear', antialias=True)
    want = rf * alpha + sf * (1 - alpha)
    got = rt.raster_level(rt.to_nhwc(tex.to(DEV)), uv.to(DEV).contiguous(), rt.to_nhwc(stat.to(DEV))[..., :Cc], (sb[0], sb[1], sb[2], sb[3]),
                          alpha[:, 0].contiguous().to(DEV), res)
    close(rt.from_nhwc(got), want, 5e-6 * max(1.0, float(want.abs().max())), f'raster_level C={Cc} r={res}')


@pytest.mark.parametrize('Cc,tr,res', [(512, 32, 32), (512, 64, 64), (256, 128, 128), (128, 256, 256)])
@pytest.mark.parametrize('uv_kind', ['smooth', 'random', 'outside'])
def test_raster_level_cell_merged_vs_per_sample(monkeypatch, Cc, tr, res, uv_kind):
    """The cell-merged kernels (raster_hpass_merge_kernel: the four texels of a cell gathered once for the run of consecutive samples of
    a row that fall into it; raster_fused_kernel: the same over an output pixel's whole 2-D window, one launch) against the per-sample kernel: the same sum in another order (fp32 reassociation only), on a smooth
    UV map (long runs), a random one (no runs at all) and one that leaves the texture (zero-padding corners)."""
    g = torch.Generator().manual_seed(Cc + res + len(uv_kind))
    B = 2
    tex = torch.randn(B, tr, tr, Cc, generator=g).to(DEV)
    uv = synth_uv(B)
    if uv_kind == 'random':
        uv = torch.cat([torch.rand(B, 256, 256, 2, generator=g) * 2 - 1, uv[..., 2:]], dim=-1)
    elif uv_kind == 'outside':
        uv = torch.cat([uv[..., :2] * 1.3, uv[..., 2:]], dim=-1)
    uv = uv.contiguous().to(DEV)
    stat = torch.randn(B, tr, tr, Cc, generator=g).to(DEV)
    sb = [round(i * tr / 256) for i in (57, 185, 64, 192)]
    alpha = torch.rand(B, res, res, generator=g).to(DEV)
    outs = {}
    for name, merge, fused in (('per_sample', '0', '0'), ('row_merge', '1', '0'), ('fused', '1', '1')):
        monkeypatch.setenv('IA_RASTER_MERGE', merge)
        monkeypatch.setenv('IA_RASTER_FUSED', fused)
        monkeypatch.setenv('IA_RASTER_FUSED_SCALE', '2')        # exercise the one-launch kernel at scale 2 as well (default: scale >= 4)
        outs[name] = rt.raster_level(tex, uv, stat, (sb[0], sb[1], sb[2], sb[3]), alpha, res).clone()
    scale = max(1.0, float(outs['per_sample'].abs().max()))
    assert float((outs['per_sample'] - outs['row_merge']).abs().max()) <= 2e-6 * scale
    # the one-launch kernel (raster_fused_kernel: a warp per output pixel, 2-D window, table of cells) sums up to 256 samples per pixel
    assert float((outs['per_sample'] - outs['fused']).abs().max()) <= 3e-6 * scale
    again = rt.raster_level(tex, uv, stat, (sb[0], sb[1], sb[2], sb[3]), alpha, res)
    assert torch.equal(again, outs['fused'])                 # deterministic: fixed sample, run and table order


def synth_uv(B):
    from invertavatar_b200 import synth
    return synth.uvcoords_image(B)


def test_layout_grid_u8():
    """Output stage (reenact_avatar_next3d.py:117-131): fused quantise + tile + HWC against the reference expression."""
    from invertavatar_b200 import glue
    g = torch.Generator().manual_seed(5)
    img = torch.randn(6, 3, 20, 12, generator=g) * 0.8
    img[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.99607843, -1.00392157])   # quantisation edges
    for gw, gh in ((6, 1), (3, 2), (2, 3)):
        ref = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8).reshape(gh, gw, 3, 20, 12).permute(2, 0, 3, 1, 4).reshape(3, gh * 20, gw * 12).permute(1, 2, 0)
        for x in (img.to(DEV), img.to(DEV).contiguous(memory_format=torch.channels_last)):
            out = glue.layout_grid(x, grid_w=gw, grid_h=gh)
            assert out.dtype == np.uint8 and out.shape == (gh * 20, gw * 12, 3)
            assert np.array_equal(out, ref.numpy())


def test_plugin_dtypes_f16_f64():
    """bias_act / upfirdn2d run natively on the element types the reference plugins dispatch (double, float, half:
    bias_act.cpp:81, upfirdn2d.cpp:67) -- float arithmetic for half, double arithmetic for double."""
    g = golden('ops.npz')
    x, b = T(g['bias_act/x']), T(g['bias_act/b'])
    for act in ('linear', 'lrelu', 'sigmoid', 'softplus', 'swish'):
        y64 = rt.bias_act(x.double().to(DEV), b.double().to(DEV), act=act)
        assert y64.dtype == torch.float64
        # (alpha / gain travel as C floats, as in the reference plugin's bias_act_kernel_params: 0.2 and sqrt(2) are rounded to fp32)
        close(y64, o_ops.bias_act(x.double(), b.double(), act=act, alpha=float(np.float32(rt.ACT_DEFAULTS[act][0])),
                                  gain=float(np.float32(rt.ACT_DEFAULTS[act][1]))), 1e-12, 'f64 ' + act)
        xh, bh = x.half(), b.half()
        y16 = rt.bias_act(xh.to(DEV), bh.to(DEV), act=act)
        assert y16.dtype == torch.float16
        want = o_ops.bias_act(xh.float(), bh.float(), act=act).half()        # float arithmetic, one rounding at the store
        close(y16.float(), want.float(), 2e-3 * max(1.0, float(want.abs().max())), 'f16 ' + act)
    xu, f = T(g['upfirdn2d/x']), T(g['upfirdn2d/f']).to(DEV)
    y64 = rt.upfirdn2d(xu.double().to(DEV), f, up=(2, 2), padding=(2, 1, 2, 1), gain=4)
    assert y64.dtype == torch.float64
    close(y64, g['upfirdn2d/up2'], 1e-6)
    y16 = rt.upfirdn2d(xu.half().to(DEV).contiguous(memory_format=torch.channels_last), f, up=(2, 2), padding=(2, 1, 2, 1), gain=4)
    assert y16.dtype == torch.float16 and y16.is_contiguous(memory_format=torch.channels_last)
    want = o_ops.upfirdn2d(xu.half().float(), T(g['upfirdn2d/f']), up=2, padding=[2, 1, 2, 1], gain=4)
    close(y16.float(), want, 4e-3)


def test_filtered_lrelu_export():
    """ia_filtered_lrelu / ia_filtered_lrelu_act (filtered_lrelu.cpp:20,217) against the reference-minted vectors, the
    library export and the bias_act/upfirdn2d composition (impl='ref'), fp32 and fp16, 1-D (separable) and 2-D filters."""
    from invertavatar_b200 import ops
    g = golden('ops.npz')
    x, b, f = T(g['filtered_lrelu/x']).to(DEV), T(g['filtered_lrelu/b']).to(DEV), T(g['filtered_lrelu/f']).to(DEV)
    for impl in ('cuda', 'ref'):
        close(ops.filtered_lrelu(x, fu=f, fd=f, b=b, up=2, down=2, padding=3, clamp=0.9, impl=impl), g['filtered_lrelu/up2_down2'], 2e-6, impl)
        close(ops.filtered_lrelu(x, b=b, impl=impl), g['filtered_lrelu/plain'], 1e-6, impl)
    f2 = torch.outer(f, f)
    close(ops.filtered_lrelu(x, fu=f2, fd=f2, b=b, up=2, down=2, padding=3, clamp=0.9), g['filtered_lrelu/up2_down2'], 2e-6, '2-D filters')
    xc = x.expand(2, -1, -1, -1).contiguous(memory_format=torch.channels_last)
    y = ops.filtered_lrelu(xc, fu=f, fd=f, b=b, up=2, down=2, padding=3, clamp=0.9)
    assert y.is_contiguous(memory_format=torch.channels_last)
    close(y[1:], g['filtered_lrelu/up2_down2'], 2e-6, 'channels-last')
    yh = ops.filtered_lrelu(x.half(), fu=f, fd=f, b=b.half(), up=2, down=2, padding=3, clamp=0.9)
    assert yh.dtype == torch.float16
    close(yh.float(), o_ops.filtered_lrelu(x.half().float().cpu(), fu=f.cpu(), fd=f.cpu(), b=b.half().float().cpu(), up=2, down=2, padding=3, clamp=0.9), 2e-3)
    xa = x.clone()
    rt.filtered_lrelu_act_(xa, gain=1.5, slope=0.1, clamp=1.0)
    close(xa, o_ops.bias_act(x.cpu(), None, act='lrelu', alpha=0.1, gain=1.5, clamp=1.0), 1e-6)


def test_integration_stub_on_golden():
    """The reference-side ctypes binding of INTEGRATION.md section 3, executed verbatim, on a reference-minted vector."""
    from types import SimpleNamespace
    from test_abi import integration_stub_namespace
    g = golden('ops.npz')
    fwd = integration_stub_namespace()['_bias_act_cuda_forward']
    x, b = T(g['bias_act/x']).to(DEV), T(g['bias_act/b']).to(DEV)
    for act in ('lrelu', 'sigmoid', 'linear'):
        spec = SimpleNamespace(cuda_idx=rt.ACT_IDS[act])
        y = fwd(x, b, 1, spec, rt.ACT_DEFAULTS[act][0], rt.ACT_DEFAULTS[act][1], None)
        close(y, g[f'bias_act/{act}'], 2e-6, act)
    y = fwd(x.half(), b.half(), 1, SimpleNamespace(cuda_idx=3), 0.2, 2 ** 0.5, 256.0)
    close(y.float(), o_ops.bias_act(x.half().float().cpu(), b.half().float().cpu(), act='lrelu', clamp=256), 4e-3)
