"""GPU parity of the improved one-shot encoder through the C-ABI: the new kernels against torch fp32 / the CPU
oracle, the transformer block, MixVisionTransformer, the two decoders and ``uvnet_new.inversionNet.forward`` against golden
vectors minted from the unmodified reference (tests/golden/make_golden_segformer.py)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import (SEG_RES, build_os_inversion_net, build_segformer_part, golden, segformer_forward_draws, segformer_inputs)
from golden.fingerprint import compare, unpack
from invertavatar_b200 import runtime as rt
from invertavatar_b200 import synth
from oracle import segformer as o_sf

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RTOL = 1e-3   # north_star tolerance: 1e-3 max-abs (relative to the tensor's magnitude where that exceeds 1)


def rel_err(got, ref):
    ref = ref.float().cpu()
    return float((got.float().cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))


def _join(sp):
    """Split (bf16 hi/lo) -> fp32 (hi + lo reproduces fp32 to 2^-16 relative)."""
    return sp.hi.float() + sp.lo.float()


def _sd(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def _cmp(t, g, key, rtol=RTOL):
    fp = unpack(key, g)
    scale = max(1.0, float(np.abs(fp['sub']).max()))
    return compare(t.float().cpu(), fp, rtol * scale, key)


@pytest.mark.parametrize('k,stride,pad,C1,C2,ps', [(7, 2, 3, 5, 8, 2), (3, 2, 1, 16, 0, 1), (4, 4, 0, 12, 0, 1), (7, 4, 3, 7, 0, 1)])
def test_im2col_bit_exact(k, stride, pad, C1, C2, ps):
    """The patch gather is data movement: bit-exact against F.unfold of the concatenated (pixel-shuffled) sources."""
    g = torch.Generator().manual_seed(k)
    H = W = 16
    x1 = torch.randn(2, C1, H, W, generator=g)
    srcs_ref = [x1]
    srcs = [x1.to(DEV).permute(0, 2, 3, 1)]
    if C2:
        x2 = torch.randn(2, C2 * ps * ps, H // ps, W // ps, generator=g)
        srcs_ref.append(F.pixel_shuffle(x2, ps))
        srcs.append((x2.to(DEV).permute(0, 2, 3, 1), ps))
    cat = torch.cat(srcs_ref, dim=1)
    Ct = cat.shape[1]
    sp = rt.enc_im2col(srcs, k, stride, pad)
    OH = (H + 2 * pad - k) // stride + 1
    cols = F.unfold(cat, k, padding=pad, stride=stride).reshape(2, Ct, k * k, OH, OH).permute(0, 3, 4, 2, 1).reshape(2, OH, OH, k * k * Ct)
    got = _join(sp).cpu()
    assert tuple(got.shape[:3]) == (2, OH, OH) and got.shape[3] % 64 == 0
    want = cols.bfloat16().float()
    want = want + (cols - want).bfloat16().float()
    assert torch.equal(got[..., :k * k * Ct], want)
    assert float(got[..., k * k * Ct:].abs().max()) == 0.0 if got.shape[3] > k * k * Ct else True


@pytest.mark.parametrize('C,eps', [(64, 1e-6), (320, 1e-6), (1024, 1e-5)])
def test_layer_norm_vs_torch(C, eps):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, 5, 7, C, generator=g) * 3 + 0.5
    pre = torch.randn(C, generator=g)
    ln = torch.nn.LayerNorm(C, eps=eps)
    with torch.no_grad():
        ln.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ln.bias.copy_(torch.randn(C, generator=g))
        want = ln(x + pre)
        sp, y = rt.layer_norm(x.to(DEV), ln.to(DEV), pre_bias=pre.to(DEV), want_split=True, want32=True)
    assert float((y.cpu() - want).abs().max()) < 5e-6 * max(1.0, float(want.abs().max()))
    assert float((_join(sp).cpu()[..., :C] - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('heads,hd,Nq_hw,Nk_hw,bias', [(4, 256, (8, 8), (8, 8), False), (4, 256, (12, 12), (12, 12), False), (1, 64, (16, 16), (2, 2), True),
                                                       (5, 64, (9, 7), (5, 3), True), (8, 64, (4, 4), (4, 4), True)])
def test_attention_vs_torch(heads, hd, Nq_hw, Nk_hw, bias):
    """Flash-style attention vs softmax(q k^T scale) v in float64: ragged tiles (N not a multiple of 64), Nk != Nq (spatial
    reduction), both instantiated head sizes, with and without q/kv bias, logits of a few units."""
    g = torch.Generator().manual_seed(heads * hd + Nq_hw[0])
    Cc = heads * hd
    B = 2
    q = torch.randn(B, *Nq_hw, Cc, generator=g)
    kv = torch.randn(B, *Nk_hw, 2 * Cc, generator=g)
    qb = torch.randn(Cc, generator=g) * 0.3 if bias else None
    kvb = torch.randn(2 * Cc, generator=g) * 0.3 if bias else None
    scale = hd ** -0.5 * 2.0
    sp, out = rt.attention(q.to(DEV), kv.to(DEV), heads, scale, None if qb is None else qb.to(DEV), None if kvb is None else kvb.to(DEV), want32=True)
    qq = (q + (qb if bias else 0)).double().reshape(B, -1, heads, hd).permute(0, 2, 1, 3)
    kk = (kv + (kvb if bias else 0)).double().reshape(B, -1, 2, heads, hd).permute(2, 0, 3, 1, 4)
    want = (((qq @ kk[0].transpose(-2, -1)) * scale).softmax(-1) @ kk[1]).transpose(1, 2).reshape(B, *Nq_hw, Cc).float()
    assert float((out.cpu() - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))
    assert float((_join(sp).cpu() - want).abs().max()) < 4e-5 * max(1.0, float(want.abs().max()))


def _split_of(x):
    """fp32 -> Split (bf16 hi/lo) on the device, as a GEMM epilogue would emit it."""
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return rt.Split(hi.contiguous(), lo.contiguous())


@pytest.mark.parametrize('heads,Nq_hw,Nk_hw,gain', [(4, (8, 8), (8, 8), 1.0), (4, (12, 12), (12, 12), 3.0), (2, (17, 9), (5, 7), 2.0), (4, (32, 32), (32, 32), 1.0)])
def test_attention_tc_vs_torch(heads, Nq_hw, Nk_hw, gain):
    """Tensor-core attention (mma.sync, 3-term bf16 split for q k^T and p v) vs float64: ragged query / key tiles (N not a multiple of
    128 / 32), Nk != Nq, logits of several units (gain), against the values the split operands actually hold."""
    g = torch.Generator().manual_seed(heads + Nq_hw[0])
    Cc = heads * 256
    B = 2
    q = torch.randn(B, *Nq_hw, Cc, generator=g).to(DEV)
    kv = torch.randn(B, *Nk_hw, 2 * Cc, generator=g).to(DEV)
    qs, kvs = _split_of(q), _split_of(kv)
    scale = 256 ** -0.5 * gain
    sp, out = rt.attention_tc(qs, kvs, heads, scale, want32=True)
    qq = _join(qs).double().reshape(B, -1, heads, 256).permute(0, 2, 1, 3)
    kk = _join(kvs).double().reshape(B, -1, 2, heads, 256).permute(2, 0, 3, 1, 4)
    want = (((qq @ kk[0].transpose(-2, -1)) * scale).softmax(-1) @ kk[1]).transpose(1, 2).reshape(B, *Nq_hw, Cc).float().cpu()
    assert float((out.cpu() - want).abs().max()) < 3e-5 * max(1.0, float(want.abs().max()))
    assert float((_join(sp).cpu() - want).abs().max()) < 5e-5 * max(1.0, float(want.abs().max()))


def test_attention_paths_agree(monkeypatch):
    """A transformer_block Block through the tensor-core attention and through the fp32 CUDA-core kernel (IA_ATTENTION=simt)."""
    tb = build_segformer_part('tb')
    blk = copy.deepcopy(tb.ViT[0]).to(DEV)
    x = torch.randn(2, 20 * 12, 1024, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        a = blk(x, 20, 12)
        monkeypatch.setenv('IA_ATTENTION', 'simt')
        b = blk(x, 20, 12)
    assert rel_err(a, b) < 5e-5


def test_dwconv_gelu_vs_torch():
    g = torch.Generator().manual_seed(3)
    C = 128
    x = torch.randn(2, 9, 11, C, generator=g)
    ib = torch.randn(C, generator=g)
    dw = torch.nn.Conv2d(C, C, 3, 1, 1, groups=C)
    with torch.no_grad():
        want = F.gelu(dw((x + ib).permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
        sp = rt.dwconv_gelu(x.to(DEV), ib.to(DEV), dw.to(DEV))
    assert float((_join(sp).cpu() - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))


def test_block_vs_oracle():
    """One Block with spatial-reduction attention and q/kv bias (MixVisionTransformer stage 2) and one of transformer_block's."""
    mit = build_segformer_part('mit')
    blk = copy.deepcopy(mit.block2[0]).to(DEV)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 8 * 8, 128, generator=g)
    with torch.no_grad():
        want = o_sf.block(_sd(blk), x, 8, 8, 2, 4, 1e-6)
        got = blk(x.to(DEV), 8, 8)
    assert rel_err(got, want) < 2e-4
    tb = build_segformer_part('tb')
    blk = copy.deepcopy(tb.ViT[1]).to(DEV)
    x = torch.randn(1, 10 * 6, 1024, generator=g)
    with torch.no_grad():
        want = o_sf.block(_sd(blk), x, 10, 6, 4, 1, 1e-5)
        got = blk(x.to(DEV), 10, 6)
    assert rel_err(got, want) < 2e-4


def test_transformer_block_golden():
    g = golden('segformer.npz')
    tb = copy.deepcopy(build_segformer_part('tb')).to(DEV)
    with torch.no_grad():
        out = tb(segformer_inputs('tb').to(DEV))
    err, _ = _cmp(out, g, 'tb/out')
    print(f'transformer_block: {err:.2e}')


def test_mix_vision_transformer_golden():
    g = golden('segformer.npz')
    mit = copy.deepcopy(build_segformer_part('mit')).to(DEV)
    with torch.no_grad():
        outs = mit(segformer_inputs('mit').to(DEV))
    assert len(outs) == 4
    for i, o in enumerate(outs):
        err, _ = _cmp(o, g, f'mit/out{i}')
        print(f'mit stage {i}: {tuple(o.shape)} {err:.2e}')


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_decoders_golden(mode):
    """TriPlanefeat_SegformerDecoder / TriPlaneSFTfeat_SegformerDecoder at their real size (256^2 input, batch 2): attention over
    up to 4096 tokens of 1024 channels; eval mode (eval_updated_os.py:93) and train-mode BatchNorm in the decoder."""
    g = golden('segformer.npz')
    for kind in ('texdec', 'tridec'):
        m = copy.deepcopy(build_segformer_part(kind)).to(DEV)
        if mode == 'train':
            m.train()
            m.input_layer.eval()
            m.body.eval()
        with torch.no_grad():
            o = m(segformer_inputs(kind).to(DEV))
        items = enumerate(o) if isinstance(o, list) else o.items()
        for k, t in items:
            err, _ = _cmp(t, g, f'{kind}/{mode}/{k}')
            print(f'{kind}/{mode}/{k}: {tuple(t.shape)} {err:.2e}')


def test_forward_golden():
    """uvnet_new.inversionNet.forward end to end as eval_updated_os.py:164-171 calls it (eval mode, return_feats=True)."""
    g = golden('segformer.npz')
    net = copy.deepcopy(build_os_inversion_net()).to(DEV)
    x, c, v = synth.encoder_inputs(1)
    x = {k: t.to(DEV) for k, t in x.items()}
    v = {k: t.to(DEV) for k, t in v.items()}
    draws = segformer_forward_draws(1)
    net.generator.renderer.depth_jitter = [d[0].to(DEV) for d in draws]
    net.generator.renderer.importance_u = [d[1].to(DEV) for d in draws]
    rt.reset_launch_count()
    with torch.no_grad():
        out = net(x, c.to(DEV), v, return_feats=True, visualize_input=True)
    assert rt.launch_count() > 0
    assert rel_err(out['w'], torch.from_numpy(g['fwd/w'])) < RTOL
    for k in ('x_input', 'e4e_image', 'image', 'image_raw', 'image_depth'):
        err, _ = _cmp(out[k], g, f'fwd/{k}')
        print(f'forward {k}: {err:.2e}')
    # feature maps of the generator's backbones carry the per-layer precision policy (see test_gpu_encoder.FEAT_RTOL)
    for i, t in enumerate(out['texture']):
        _cmp(t, g, f'fwd/texture{i}', 3e-3)
    for i, t in enumerate(out['static']):
        _cmp(t, g, f'fwd/static{i}', 3e-3)
    with pytest.raises(AttributeError):      # the reference class has no AR_eval_forward (uvnet_new.py)
        net.AR_eval_forward()
