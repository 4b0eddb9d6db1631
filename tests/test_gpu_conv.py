"""GPU parity of the modulated-convolution path: tcgen05 implicit GEMM (ia_conv_tc) and the CUDA-core cross-check
(ia_conv_simt) against the oracle's modulated_conv2d / synthesis_layer / torgb_layer / synthesis_block."""
import numpy as np
import pytest
import torch

from common import T, golden
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import stylegan2 as sg
from oracle import stylegan2 as o_sg

pytestmark = pytest.mark.gpu
DEV = 'cuda'
# 3-term bf16 split (hi*hi + hi*lo + lo*hi, fp32 accumulate): operands carry 16 mantissa bits, the dropped lo*lo term is
# 2^-16 relative per product -> a few 1e-5 relative to the output scale (SURVEY 8d precision budget: 5.6e-5 on the image)
TOL = 5e-5


def maxerr(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.fixture(params=['simt', 'tc'])
def impl(request):
    old = rt.get_conv_impl()
    rt.set_conv_impl(request.param)
    yield request.param
    rt.set_conv_impl(old)


def _layer(cin, cout, res, up, seed, w_dim=64):
    torch.manual_seed(seed)
    L = sg.SynthesisLayer(cin, cout, w_dim=w_dim, resolution=res, up=up).requires_grad_(False)
    L.noise_strength.fill_(0.3)
    L.bias.copy_(torch.randn(cout) * 0.2)
    return L


@pytest.mark.parametrize('cin,cout,res,up,B', [
    (6, 4, 8, 1, 2), (6, 4, 16, 2, 2),            # tiny, heavy padding
    (64, 64, 4, 1, 3), (128, 96, 8, 2, 1),        # b4/b8-like
    (512, 512, 16, 1, 2), (512, 256, 32, 2, 2),   # full-width layers
    (256, 128, 64, 2, 1), (128, 128, 64, 1, 2),   # b256-like at reduced resolution
    (32, 256, 32, 2, 2),                          # SR block0 conv0 (Cin 32 -> padded to 64)
    (40, 72, 12, 1, 2), (40, 72, 24, 2, 5),       # odd sizes: ragged tiles, Cout not a multiple of 32
])
def test_synthesis_layer(impl, cin, cout, res, up, B):
    L = _layer(cin, cout, res, up, seed=cin + cout)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, cin, res // up, res // up, generator=g)
    w = torch.randn(B, 64, generator=g)
    ref = o_sg.synthesis_layer(L.state_dict(), x, w, up=up, noise_mode='const', gain=0.8, conv_clamp=3.0)
    L.conv_clamp = 3.0
    y = L.to(DEV)(x.to(DEV), w.to(DEV), noise_mode='const', gain=0.8)
    assert tuple(y.shape) == tuple(ref.shape)
    scale = float(ref.abs().max())
    err = maxerr(y, ref)
    assert err <= TOL * max(scale, 1.0), f'{impl}: err {err:.3e} (scale {scale:.2f})'


def test_synthesis_layer_noise_modes(impl):
    L = _layer(16, 8, 8, 1, seed=1)
    g = torch.Generator().manual_seed(8)
    x, w = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 64, generator=g)
    ref = o_sg.synthesis_layer(L.state_dict(), x, w, noise_mode='none')
    y = L.to(DEV)(x.to(DEV), w.to(DEV), noise_mode='none')
    assert maxerr(y, ref) <= TOL * max(1.0, float(ref.abs().max()))
    y1 = L(x.to(DEV), w.to(DEV), noise_mode='random')
    y2 = L(x.to(DEV), w.to(DEV), noise_mode='random')
    assert maxerr(y1, y2) > 1e-3     # fresh noise each call


@pytest.mark.parametrize('cin,cimg', [(6, 3), (512, 96), (128, 32), (256, 3)])
def test_torgb(impl, cin, cimg):
    torch.manual_seed(cin)
    L = sg.ToRGBLayer(cin, cimg, w_dim=64, conv_clamp=2.0).requires_grad_(False)
    L.bias.copy_(torch.randn(cimg) * 0.2)
    g = torch.Generator().manual_seed(9)
    x, w = torch.randn(2, cin, 16, 16, generator=g), torch.randn(2, 64, generator=g)
    ref = o_sg.torgb_layer(L.state_dict(), x, w, conv_clamp=2.0)
    y = L.to(DEV)(x.to(DEV), w.to(DEV))
    assert maxerr(y, ref) <= TOL * max(1.0, float(ref.abs().max()))


def test_modconv_golden(impl):
    """Reference-minted vectors (ops.npz) through the layer modules."""
    g = golden('ops.npz')
    x, w, s = T(g['modconv/x']), T(g['modconv/w']), T(g['modconv/s'])
    # build a layer whose affine produces exactly the golden styles: affine.weight = 0, affine.bias = s (per sample -> B=1 each)
    for b in range(2):
        for up, key, nz in ((1, 'modconv/same', 'modconv/noise'), (2, 'modconv/up2', 'modconv/noise2')):
            L = sg.SynthesisLayer(6, 4, w_dim=8, resolution=8 * up, up=up, activation='linear').requires_grad_(False)
            L.weight.copy_(w)
            L.affine.weight.zero_()
            L.affine.bias.copy_(s[b])
            L.noise_const.copy_(T(g[nz]))
            L.noise_strength.fill_(1.0)
            y = L.to(DEV)(x[b:b + 1].to(DEV), torch.zeros(1, 8, device=DEV), noise_mode='const', gain=1)
            assert maxerr(y, T(g[key])[b:b + 1]) <= TOL * max(1.0, float(T(g[key]).abs().max())), (key, b)


def test_synthesis_block_skip(impl):
    """conv0(up2)+conv1+ToRGB+skip-image upsample, as one block (networks_stylegan2_new.py:417-467)."""
    torch.manual_seed(3)
    blk = sg.SynthesisBlock(64, 32, w_dim=64, resolution=16, img_channels=8, is_last=False, conv_clamp=None, use_fp16=False).requires_grad_(False)
    for n, p in blk.named_parameters():
        if n.endswith('noise_strength'):
            p.fill_(0.2)
        if n.endswith('.bias') and 'affine' not in n:
            p.copy_(torch.randn_like(p) * 0.1)
    g = torch.Generator().manual_seed(4)
    x, img, ws = torch.randn(2, 64, 8, 8, generator=g), torch.randn(2, 8, 8, 8, generator=g), torch.randn(2, 3, 64, generator=g)
    rx, rimg = o_sg.synthesis_block(blk.state_dict(), x, img, ws, noise_mode='const')
    blk = blk.to(DEV)
    yx, yimg = blk(x.to(DEV), img.to(DEV), ws.to(DEV), noise_mode='const')
    assert maxerr(yx, rx) <= TOL * max(1.0, float(rx.abs().max())) and maxerr(yimg, rimg) <= TOL * max(1.0, float(rimg.abs().max()))
    # CS-SFT condition on the second half of the channels (:448-452)
    cond = torch.stack([torch.randn(2, 16, 16, 16, generator=g), torch.randn(2, 16, 16, 16, generator=g)])
    rx, rimg = o_sg.synthesis_block(blk.cpu().state_dict(), x, img, ws, condition=cond, noise_mode='const')
    yx, yimg = blk.to(DEV)(x.to(DEV), img.to(DEV), ws.to(DEV), condition=cond.to(DEV), noise_mode='const')
    assert maxerr(yx, rx) <= TOL * max(1.0, float(rx.abs().max())) and maxerr(yimg, rimg) <= TOL * max(1.0, float(rimg.abs().max()))


def test_tc_matches_simt_large():
    """Full-size b64 layer (512->512 @64^2, batch 2): tensor-core result against the CUDA-core path (same operands)."""
    L = _layer(512, 512, 64, 1, seed=11).to(DEV)
    x = torch.randn(2, 512, 64, 64, device=DEV)
    w = torch.randn(2, 64, device=DEV)
    rt.set_conv_impl('simt')
    a = L(x, w, noise_mode='const')
    rt.set_conv_impl('tc')
    b = L(x, w, noise_mode='const')
    assert maxerr(a, b) <= TOL * max(1.0, float(a.abs().max()))


def _small_network(seed=5):
    torch.manual_seed(seed)
    net = sg.SynthesisNetwork(w_dim=64, img_resolution=32, img_channels=8, channel_base=1280, channel_max=40, num_fp16_res=0,
                              conv_clamp=None).requires_grad_(False)
    for n, p in net.named_parameters():
        if n.endswith('noise_strength'):
            p.fill_(0.2)
        if n.endswith('.bias') and 'affine' not in n:
            p.copy_(torch.randn_like(p) * 0.1)
    return net


def test_synthesis_network_fused_chain(impl):
    """Whole-network fused chain (epilogues emit the next layer's operands) with channel counts that need zero padding
    (40 -> 64), return_list and cond_list semantics of networks_stylegan2_new.py:509-548."""
    net = _small_network()
    g = torch.Generator().manual_seed(6)
    ws = torch.randn(3, net.num_ws, 64, generator=g)
    sd = net.state_dict()
    ref = o_sg.synthesis_network(sd, ws, return_list=True, out_res=(8, 32))
    net = net.to(DEV)
    got = net(ws.to(DEV), return_list=True, out_res=(8, 32), noise_mode='const')
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert maxerr(a, b) <= TOL * max(1.0, float(b.abs().max()))
    # cond_list: img blended at res 8, x blended at res 8 and 16 (index < end_layer)
    conds = [torch.cat([torch.randn(3, c, r, r, generator=g), torch.rand(3, 1, r, r, generator=g)], 1) for (c, r) in ((8, 8), (40, 8), (40, 16))]
    ref = o_sg.synthesis_network(sd, ws, cond_list=conds, return_list=False, out_res=(8, 32))
    got = net(ws.to(DEV), cond_list=[c.to(DEV) for c in conds], return_list=False, out_res=(8, 32), noise_mode='const')
    assert maxerr(got, ref) <= TOL * max(1.0, float(ref.abs().max()))


def test_fused_torgb_matches_oracle(monkeypatch):
    """3-channel ToRGB contracted inside conv1's epilogue (ia_emit.rgb_*; 256 channels = two N tiles adding into the same
    pixels) against the oracle, and bit-for-bit reproducible; IA_FUSE_TORGB=0 (separate 1x1 convolution) agrees to tolerance."""
    torch.manual_seed(9)
    net = sg.SynthesisNetwork(w_dim=64, img_resolution=32, img_channels=3, channel_base=8192, channel_max=256, num_fp16_res=0,
                              conv_clamp=None).requires_grad_(False)
    for n, p in net.named_parameters():
        if n.endswith('noise_strength'):
            p.fill_(0.2)
        if n.endswith('.bias') and 'affine' not in n:
            p.copy_(torch.randn_like(p) * 0.1)
    g = torch.Generator().manual_seed(3)
    ws = torch.randn(2, net.num_ws, 64, generator=g)
    ref = o_sg.synthesis_network(net.state_dict(), ws)
    net = net.to(DEV)
    assert rt.can_fuse_torgb(32, 32, 256, 3) and not rt.can_fuse_torgb(8, 8, 256, 3)
    a = net(ws.to(DEV), noise_mode='const')
    b = net(ws.to(DEV), noise_mode='const')
    assert torch.equal(a, b)
    assert maxerr(a, ref) <= TOL * max(1.0, float(ref.abs().max()))
    monkeypatch.setenv('IA_FUSE_TORGB', '0')
    c = net(ws.to(DEV), noise_mode='const')
    assert maxerr(c, ref) <= TOL * max(1.0, float(ref.abs().max()))


def test_fused_torgb_tail_bit_identical(monkeypatch):
    """ToRGB tail inside the 1x1 convolution's epilogue (mode 2) == 1x1 convolution + ia_torgb_finish, bit for bit (with and
    without a previous image, 32 and 96 image channels)."""
    for cin, cimg, res, with_prev in ((128, 32, 32, True), (256, 96, 16, True), (128, 32, 16, False)):
        torch.manual_seed(cin + cimg)
        L = sg.ToRGBLayer(cin, cimg, w_dim=64, conv_clamp=0.8).requires_grad_(False)
        L.bias.copy_(torch.randn(cimg) * 0.3)
        L = L.to(DEV)
        B = 3
        x = torch.randn(B, res, res, cin, device=DEV)
        st = torch.randn(B, cin, device=DEV)
        prev = torch.randn(B, res // 2, res // 2, cimg, device=DEV) if with_prev else None
        hi, lo = rt.modsplit(x, st, C_pad=L.pack().Cin_pad)
        outs = []
        for flag in ('1', '0'):
            monkeypatch.setenv('IA_FUSE_TORGB_TAIL', flag)
            assert rt.can_fuse_torgb_tail(res, res, cimg) == (flag == '1')
            outs.append(L.run_split(rt.Split(hi, lo), img_prev=prev).clone())
        assert torch.equal(outs[0], outs[1]), (cin, cimg, res, with_prev)


def test_row_padded_operand_layout_bit_identical(monkeypatch):
    """Transposed convolutions fed with the row-padded operand layout ([B][H+1][W][C], zero row after every image: tiles of the
    phase grids run across image boundaries) give bit-identical results to the dense layout -- producers: conv epilogue (emit 1
    image stride), grouped-free chain, modsplit with blend; 3 images so that tiles straddle two images."""
    torch.manual_seed(21)
    net = sg.SynthesisNetwork(w_dim=64, img_resolution=128, img_channels=8, channel_base=2560, channel_max=40, num_fp16_res=0,
                              conv_clamp=None).requires_grad_(False)
    for n, p in net.named_parameters():
        if n.endswith('noise_strength'):
            p.fill_(0.2)
    g = torch.Generator().manual_seed(4)
    ws = torch.randn(3, net.num_ws, 64, generator=g).to(DEV)
    conds = [torch.cat([torch.randn(3, c, r, r, generator=g), torch.rand(3, 1, r, r, generator=g)], 1).to(DEV)
             for (c, r) in ((8, 32), (40, 32), (40, 64))]
    net = net.to(DEV)
    outs = {}
    for flag in ('1', '0'):
        monkeypatch.setenv('IA_CONV_CAT_ROWS', flag)
        assert rt.pad_row_wanted(32, 32) == (flag == '1') and not rt.pad_row_wanted(16, 16)
        a = net(ws, return_list=True, out_res=(32, 128), noise_mode='const')
        b = net(ws, cond_list=conds, return_list=False, out_res=(32, 128), noise_mode='const')
        outs[flag] = [t.clone() for t in a] + [b.clone()]
    for x, y in zip(outs['1'], outs['0']):
        assert torch.equal(x, y)


@pytest.mark.parametrize('B,H,cin,cout,k', [(4, 16, 1024, 512, 3), (4, 32, 256, 256, 3), (1, 16, 512, 512, 3), (2, 32, 384, 96, 1), (3, 24, 200, 72, 3)])
def test_split_k_matches_unsplit_and_is_deterministic(monkeypatch, B, H, cin, cout, k):
    """Low-resolution / small-batch convolutions (the encoder UNets, the <= 32^2 generator layers) split their input channels
    over several CTAs per tile; the partial accumulators are summed in split order by the last CTA to arrive, so the result
    (a) equals the unsplit kernel up to fp32 reassociation, (b) matches fp32 conv2d on the CPU and (c) is bit-identical from
    run to run."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(B * 100 + H + cin)
    x = torch.randn(B, cin, H, H, generator=g)
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2, bias=False).requires_grad_(False)
    conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / np.sqrt(cin * k * k))
    want = F.conv2d(x, conv.weight, padding=k // 2).permute(0, 2, 3, 1)
    conv = conv.to(DEV)
    xn = x.to(DEV).permute(0, 2, 3, 1)
    pack = rt.ConvPack.current(conv, '_ia_pack', conv.weight, need_wsq=False)

    def run():
        a, _ = rt.enc_prep([xn], C_pad=pack.Cin_pad)
        return rt.enc_conv(a, conv).clone()
    monkeypatch.setenv('IA_CONV_SPLITK', '0')
    plain = run()
    monkeypatch.delenv('IA_CONV_SPLITK')
    s1 = run()
    s2 = run()
    torch.cuda.synchronize()
    scale = max(1.0, float(want.abs().max()))
    assert maxerr(plain, want) <= TOL * scale
    assert maxerr(s1, want) <= TOL * scale, maxerr(s1, want)
    assert maxerr(s1, plain) <= TOL * scale       # (fp32 accumulation order differs: K = 9 * cin products per output)
    assert torch.equal(s1, s2), 'split-K sum order must not depend on the arrival order of the CTAs'
    # the ticket counters are left at zero for the next launch
    for ws, cnt in rt._SPLITK.values():
        assert int(cnt.abs().max()) == 0


@pytest.mark.parametrize('cin,cout,res,up,B', [(512, 512, 16, 1, 2), (512, 256, 32, 2, 2), (128, 128, 64, 1, 2), (256, 128, 64, 2, 1),
                                              (40, 72, 24, 2, 5), (64, 64, 4, 1, 3), (128, 96, 8, 2, 1)])
def test_synthesis_layer_single_pass_fp16(cin, cout, res, up, B):
    """IA_OPFMT_F16X1 (the format TriPlaneGenerator assigns to its backbone layers): one fp16 MMA per k-step.  Against the fp32
    oracle the result carries the operand rounding (2^-11 relative per element: ~1e-3 of the output scale); against the
    CUDA-core kernel reading the SAME fp16 operands it is tight (only the accumulation order differs)."""
    L = _layer(cin, cout, res, up, seed=cin + cout)
    L.tc_fmt = rt.FMT_F16X1
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, cin, res // up, res // up, generator=g)
    w = torch.randn(B, 64, generator=g)
    ref = o_sg.synthesis_layer(L.state_dict(), x, w, up=up, noise_mode='const', gain=0.8, conv_clamp=3.0)
    L.conv_clamp = 3.0
    L = L.to(DEV)
    old = rt.get_conv_impl()
    try:
        rt.set_conv_impl('tc')
        y = L(x.to(DEV), w.to(DEV), noise_mode='const', gain=0.8)
        assert L.pack().fmt == rt.FMT_F16X1 and L.pack().w_lo is None and L.pack().w_hi.dtype == torch.float16
        rt.set_conv_impl('simt')
        y_simt = L(x.to(DEV), w.to(DEV), noise_mode='const', gain=0.8)
    finally:
        rt.set_conv_impl(old)
    scale = max(1.0, float(ref.abs().max()))
    assert maxerr(y, ref) <= 2e-3 * scale, maxerr(y, ref)
    assert maxerr(y, y_simt) <= TOL * scale, maxerr(y, y_simt)


def test_synthesis_network_mixed_precision_chain(monkeypatch):
    """Fused chain with per-layer formats: 3x3 layers single-pass fp16, ToRGB layers 3-term (every epilogue emits its consumer's
    operand in the consumer's format); IA_CONV_PRECISION=bf16x3 restores the strict path on the same module."""
    net = _small_network()
    for m in net.modules():
        if isinstance(m, sg.SynthesisLayer):
            m.tc_fmt = rt.FMT_F16X1
    g = torch.Generator().manual_seed(3)
    ws = torch.randn(2, net.num_ws, 64, generator=g)
    ref = o_sg.synthesis_network(net.state_dict(), ws, return_list=True, out_res=(8, 32))
    net = net.to(DEV)
    got = net(ws.to(DEV), return_list=True, out_res=(8, 32), noise_mode='const')
    monkeypatch.setenv('IA_CONV_PRECISION', 'bf16x3')
    strict = net(ws.to(DEV), return_list=True, out_res=(8, 32), noise_mode='const')
    monkeypatch.delenv('IA_CONV_PRECISION')
    assert len(got) == len(ref)
    worst = 0.0
    for a, s_, r in zip(got, strict, ref):
        scale = max(1.0, float(r.abs().max()))
        assert maxerr(s_, r) <= 2 * TOL * scale
        e = maxerr(a, r) / scale
        worst = max(worst, e)
        assert e <= 5e-3, e
    assert worst > 1e-5, 'the single-pass format was not exercised'


@pytest.mark.parametrize('fmt', ['bf16x3', 'f16x1'])
@pytest.mark.parametrize('cin,cout,res,up,B', [
    (128, 128, 64, 1, 3),      # 48 M tiles x 1 N tile
    (256, 256, 32, 1, 5),      # 20 M tiles x 2 N tiles
    (128, 96, 48, 1, 3),       # ragged tiles, odd tile count per N tile -> a padding tile in the last pair, Cout not a multiple of 32
    (256, 128, 64, 2, 3),      # merged transposed-conv phases: (H+1)^2 grids, per-phase pair lists padded to even
    (64, 64, 40, 2, 1),        # phases with an odd number of tiles in every list
])
def test_cta_pair_matches_single_cta(monkeypatch, fmt, cin, cout, res, up, B):
    """tcgen05 cta_group::2 launches (M = 256 MMAs over two CTAs' pixel tiles, weight tile split across the pair) accumulate every
    output element over k in the same order as the 1-CTA kernel: bit-identical layer outputs, in both operand formats."""
    old = rt.get_conv_impl()
    rt.set_conv_impl('tc')
    try:
        L = _layer(cin, cout, res, up, seed=cin + cout + res).to(DEV)
        L.tc_fmt = rt.FMT_F16X1 if fmt == 'f16x1' else rt.FMT_BF16X3
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, cin, res // up, res // up, generator=g).to(DEV)
        w = torch.randn(B, 64, generator=g).to(DEV)
        monkeypatch.setenv('IA_CONV_SPLITK', '0')            # (a split launch is never a pair launch)
        monkeypatch.setenv('IA_CONV_PAIR', '0')
        y0 = L(x, w, noise_mode='const', gain=1.0).clone()
        monkeypatch.setenv('IA_CONV_PAIR', '1')
        monkeypatch.setenv('IA_CONV_PAIR_MIN_TILES', '0')    # pair launches whatever the grid size
        y1 = L(x, w, noise_mode='const', gain=1.0).clone()
        assert torch.isfinite(y1).all()
        assert torch.equal(y0, y1), float((y0 - y1).abs().max())
    finally:
        rt.set_conv_impl(old)
