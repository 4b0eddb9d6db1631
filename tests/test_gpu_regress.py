"""Regressions found on the GPU."""
import os

import pytest
import torch

from common import build_generator
from invertavatar_b200 import synth

pytestmark = pytest.mark.gpu


def test_expanded_ws_batch():
    """ws.expand(T, -1, -1) (stride-0 batch, uvnet.py:176) must give every frame the same styles as a materialised copy."""
    import copy
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    ws1 = G.mapping(synth.latents(1).cuda(), synth.frontal_camera(1).cuda(), truncation_psi=0.7, truncation_cutoff=14)
    with torch.no_grad():
        a = G.texture_backbone.synthesis(ws1.expand(3, -1, -1), cond_list=None, return_list=True, noise_mode='const')
        b = G.texture_backbone.synthesis(ws1.expand(3, -1, -1).contiguous(), cond_list=None, return_list=True, noise_mode='const')
    for x, y in zip(a, b):
        assert torch.equal(x, y)
        assert torch.equal(x[0], x[2])


@pytest.mark.parametrize('precision,tol', [('bf16x3', 1e-4), ('auto', 2e-3)])
def test_grouped_prefix_matches_sequential(monkeypatch, precision, tol):
    """The grouped evaluation of the three backbones' low-resolution blocks (ia_conv_params.groups) must reproduce the
    one-network-at-a-time path: same kernels, same operands, different tiling only -- i.e. a different fp32 accumulation
    order, whose last-bit differences the next layer's operand rounding turns into (rare) one-ulp operand differences:
    2^-16 relative in the strict 3-term format, 2^-11 in the single-pass fp16 format of the shipped policy."""
    import copy
    monkeypatch.setenv('IA_CONV_PRECISION', precision)
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    B = 3
    z, cond, c, uv = synth.latents(B).cuda(), synth.frontal_camera(B).cuda(), synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()
    jit = synth.depth_jitter(B, 64 * 64, 16).cuda()
    outs = {}
    with torch.no_grad():
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        for flag in ('1', '0'):
            monkeypatch.setenv('IA_GROUPED_PREFIX', flag)
            outs[flag] = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True,
                                     depth_jitter=jit.clone(), return_featmap=True)
    for k in ('image', 'image_raw', 'feature_image', 'triplane'):
        d = float((outs['1'][k].float() - outs['0'][k].float()).abs().max())
        assert d <= tol * max(1.0, float(outs['0'][k].abs().max())), (k, d)
    for a, b in zip(outs['1']['texture'], outs['0']['texture']):
        assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_graphed_synthesis_matches_eager():
    """Whole-frame CUDA graph (invertavatar_b200.graphs) of the batch-1 per-frame call == the eager call, also after the
    camera and mesh condition are swapped between replays."""
    import copy
    from invertavatar_b200.graphs import GraphedSynthesis
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    z, cond = synth.latents(1).cuda(), synth.frontal_camera(1).cuda()
    cams, uvs = synth.cameras(3).cuda(), synth.uvcoords_image(3).cuda()
    with torch.no_grad():
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        jit = synth.depth_jitter(1, 64 * 64, 16).cuda()
        G.renderer.fixed_jitter = jit            # deterministic depth jitter for the comparison (see rendering.py)
        gs = GraphedSynthesis(G, ws, cams[:1], uvs[:1], neural_rendering_resolution=64)
        for i in (1, 2, 0):
            got = gs(cams[i:i + 1], uvs[i:i + 1])['image'].clone()
            want = G.synthesis(ws, cams[i:i + 1], {'uvcoords_image': uvs[i:i + 1]}, neural_rendering_resolution=64, noise_mode='const',
                               evaluation=True)['image']
            assert float((got - want).abs().max()) <= 1e-6, i
        G.renderer.fixed_jitter = None


def test_two_graphs_over_one_module_replay_concurrently():
    """Two GraphedSynthesis objects over the SAME generator, with different latents, replayed at the same time on two streams: each
    must reproduce its own eager result.  The engine's persistent state (split-K workspace / tickets, the static style and
    demodulation buffers of a StylePlan) is private to a captured graph (runtime.capture_scope); shared, the second replay would
    overwrite the first one's styles or corrupt its split-K tickets."""
    import copy
    from invertavatar_b200.graphs import GraphedSynthesis
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    cond = synth.frontal_camera(1).cuda()
    cams, uvs = synth.cameras(2).cuda(), synth.uvcoords_image(2).cuda()
    with torch.no_grad():
        G.renderer.fixed_jitter = synth.depth_jitter(1, 64 * 64, 16).cuda()
        wss = [G.mapping(synth.latents(1, first=k).cuda(), cond, truncation_psi=0.7, truncation_cutoff=14) for k in (0, 5)]
        assert float((wss[0] - wss[1]).abs().max()) > 1e-2
        want = [G.synthesis(wss[k], cams[k:k + 1], {'uvcoords_image': uvs[k:k + 1]}, neural_rendering_resolution=64, noise_mode='const',
                            evaluation=True)['image'].clone() for k in range(2)]
        gs = [GraphedSynthesis(G, wss[k], cams[k:k + 1], uvs[k:k + 1], neural_rendering_resolution=64) for k in range(2)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        torch.cuda.synchronize()
        for rep in range(4):
            outs = []
            for k in range(2):
                streams[k].wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(streams[k]):
                    outs.append(gs[k]()['image'])
            for s_ in streams:
                torch.cuda.current_stream().wait_stream(s_)
            for k in range(2):
                assert float((outs[k] - want[k]).abs().max()) <= 1e-6, (rep, k)
        G.renderer.fixed_jitter = None


@pytest.mark.parametrize('B,OH,C,use32,use1,use2', [(2, 128, 64, True, True, True), (1, 64, 128, False, True, False), (3, 32, 32, True, False, True)])
def test_fir_epilogue_variants_agree(monkeypatch, B, OH, C, use32, use1, use2):
    """The three FIR-epilogue kernels (TMA-fed ring, two-column register kernel, one-column register kernel) perform the same
    arithmetic in the same order: bit-identical outputs (fp32 copy and both emitted bf16 hi/lo operands)."""
    from invertavatar_b200 import runtime as rt
    g = torch.Generator().manual_seed(B * 1000 + OH + C)
    raw = torch.randn(B, OH + 1, OH + 1, C, generator=g).cuda()
    dcoef, bias = torch.rand(B, C, generator=g).cuda() + 0.5, torch.randn(C, generator=g).cuda()
    noise, strength = torch.randn(OH, OH, generator=g).cuda(), torch.tensor(0.3).cuda()
    s1, s2 = torch.randn(B, C, generator=g).cuda(), torch.randn(B, C, generator=g).cuda()
    outs = {}
    variants = [('tma', {'IA_FIR_TMA': '1'}), ('x2', {'IA_FIR_TMA': '0', 'IA_FIR_X2': '1'}), ('x1', {'IA_FIR_TMA': '0', 'IA_FIR_X2': '0'}),
                ('x2_ring', {'IA_FIR_TMA': '0', 'IA_FIR_X2': '1', 'IA_FIR_NOISE_PREFETCH': '1', 'IA_FIR_RING': '1'})]
    if os.environ.get('IA_TEST_OPTIN'):      # opt-in code paths that have not been through the suite on hardware yet
        variants += [('x2_npf', {'IA_FIR_TMA': '0', 'IA_FIR_X2': '1', 'IA_FIR_NOISE_PREFETCH': '1'}),
                     ('tma_npf', {'IA_FIR_TMA': '1', 'IA_FIR_NOISE_PREFETCH': '1'})]
    for name, env in variants:
        monkeypatch.setenv('IA_FIR_NOISE_PREFETCH', '0')
        monkeypatch.setenv('IA_FIR_RING', '0')
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = torch.zeros(B, OH, OH, C, device='cuda') if use32 else None
        e1 = rt.new_split(B, OH, OH, max(64, C), 'cuda', C=C) if use1 else None
        e2 = rt.new_split(B, OH, OH, max(64, C), 'cuda', C=C) if use2 else None
        rt.fir_epilogue(raw, rt.fir4x4_gain4('cuda'), out, dcoef, noise, strength, bias, 'lrelu', 1.3, 2.0,
                        e1=(e1, s1) if use1 else None, e2=(e2, s2) if use2 else None)
        torch.cuda.synchronize()
        outs[name] = [t.clone() for t in ([out] if use32 else []) + ([e1.hi, e1.lo] if use1 else []) + ([e2.hi, e2.lo] if use2 else [])]
    for name in outs:
        for a, b in zip(outs['tma'], outs[name]):
            assert torch.equal(a, b), name


def test_prepack_builds_every_pack_once():
    """invertavatar_b200.prepack packs all conv / ToRGB weights at load; the first synthesis afterwards reuses them."""
    import copy
    import invertavatar_b200
    from invertavatar_b200 import runtime as rt
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    nbytes = invertavatar_b200.prepack(G)
    assert nbytes > 150e6                       # 88 M parameters: fp16 for the backbone 3x3 layers (2 B), bf16 hi/lo pairs (4 B) elsewhere
    packs = {id(m): m.__dict__['_ia_pack'] for m in G.modules() if '_ia_pack' in m.__dict__}
    assert len(packs) == 3 * 20 + 6             # per backbone: b4 conv1+torgb, b8..b256 conv0+conv1+torgb; SR: 2 blocks x 3
    z, cond, c, uv = synth.latents(1).cuda(), synth.frontal_camera(1).cuda(), synth.cameras(1).cuda(), synth.uvcoords_image(1).cuda()
    with torch.no_grad():
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True)
    for m in G.modules():
        if id(m) in packs:
            assert m.__dict__['_ia_pack'] is packs[id(m)]


def test_single_pass_decoder_within_north_star(monkeypatch):
    """Decoder MLP as single-pass fp16 MMAs (the generator's default) against the 3-term path (strict mode): the final image stays
    well inside the north-star tolerance (CPU probe: 1.2e-4 max-abs, tools/probe_render_precision.py)."""
    import copy
    from invertavatar_b200 import runtime as rt
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    z, cond, c, uv = synth.latents(2).cuda(), synth.frontal_camera(2).cuda(), synth.cameras(2).cuda(), synth.uvcoords_image(2).cuda()
    jit = synth.depth_jitter(2, 64 * 64, 16).cuda()
    outs = {}
    for m in G.modules():                                     # isolate the decoder: convolutions 3-term in both runs
        if hasattr(m, 'tc_fmt'):
            m.tc_fmt = rt.FMT_BF16X3
    with torch.no_grad():
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        for mode in ('fp32x3', 'fp16'):
            G.renderer.mlp_fmt = rt.FMT_F16X1 if mode == 'fp16' else rt.FMT_BF16X3
            outs[mode] = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True,
                                     depth_jitter=jit)['image'].clone()
    err = float((outs['fp16'] - outs['fp32x3']).abs().max())
    assert 0 < err <= 5e-4, err


def test_dead_code_elimination_is_bit_identical(monkeypatch):
    """TriPlaneGenerator.synthesis does not launch the texture backbone's 256^2 block nor its skip images after img32 (nothing
    reads them unless return_featmap): the outputs are bit-identical to the full evaluation, with fewer launches."""
    import copy
    from invertavatar_b200 import runtime as rt
    G = copy.deepcopy(build_generator(16, 16)).to('cuda')
    z, cond, c, uv = synth.latents(2).cuda(), synth.frontal_camera(2).cuda(), synth.cameras(2).cuda(), synth.uvcoords_image(2).cuda()
    jit = synth.depth_jitter(2, 64 * 64, 16).cuda()
    outs, launches = {}, {}
    with torch.no_grad():
        ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
        for flag in ('0', '1'):
            monkeypatch.setenv('IA_PRUNE_DEAD', flag)
            rt.reset_launch_count()
            outs[flag] = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True,
                                     depth_jitter=jit.clone())
            launches[flag] = rt.launch_count()
        full = G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True,
                           depth_jitter=jit.clone(), return_featmap=True)
    for k in ('image', 'image_raw', 'image_depth'):
        assert torch.equal(outs['0'][k], outs['1'][k]), k
        # (return_featmap hands the tri-planes back in fp32 and renders from them; the default call renders from fp16 planes)
        assert float((full[k] - outs['1'][k]).abs().max()) <= 1e-4, k
    assert len(full['texture']) == 6 and launches['1'] < launches['0']
    assert tuple(full['triplane'].shape) == (2, 3, 32, 256, 256) and full['triplane'].dtype == torch.float32


def test_pack_cache_round_trip(tmp_path):
    """prepack(G, source_hash=...) writes the packed weights once and reads them back on the next load: no packing kernels, the
    same tensors bit for bit, the same frame."""
    import copy
    import invertavatar_b200
    from invertavatar_b200 import runtime as rt
    G1 = copy.deepcopy(build_generator(16, 16)).to('cuda')
    n1 = invertavatar_b200.prepack(G1, source_hash='test-checkpoint', cache_dir=str(tmp_path))
    path = rt.pack_cache_path('test-checkpoint', cache_dir=str(tmp_path))
    assert os.path.exists(path) and os.path.getsize(path) > 100e6
    G2 = copy.deepcopy(build_generator(16, 16)).to('cuda')
    torch.cuda.synchronize()
    rt.reset_launch_count()
    n2 = invertavatar_b200.prepack(G2, source_hash='test-checkpoint', cache_dir=str(tmp_path))
    assert n2 == n1 and rt.launch_count() == 0, 'a cached load must not run the packing kernels'
    for (na, ma), (nb, mb) in zip(G1.named_modules(), G2.named_modules()):
        pa, pb = ma.__dict__.get('_ia_pack'), mb.__dict__.get('_ia_pack')
        assert (pa is None) == (pb is None)
        if pa is not None:
            assert pa.fmt == pb.fmt and torch.equal(pa.w_hi, pb.w_hi) and (pa.w_lo is None or torch.equal(pa.w_lo, pb.w_lo))
            assert pa.wsq is None or torch.equal(pa.wsq, pb.wsq)
    z, cond, c, uv = synth.latents(1).cuda(), synth.frontal_camera(1).cuda(), synth.cameras(1).cuda(), synth.uvcoords_image(1).cuda()
    jit = synth.depth_jitter(1, 64 * 64, 16).cuda()
    with torch.no_grad():
        imgs = []
        for G in (G1, G2):
            ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
            imgs.append(G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=64, noise_mode='const', evaluation=True,
                                    depth_jitter=jit.clone())['image'])
    assert torch.equal(imgs[0], imgs[1])
    for m in G2.modules():          # the cached packs are the ones the forward used (keys match the parameters)
        if '_ia_pack' in m.__dict__ and hasattr(m, 'pack'):
            assert m.pack() is m.__dict__['_ia_pack']


def test_graphed_call_matches_eager():
    """graphs.GraphedCall (whole call sequence in one CUDA graph) on the e4e encoder: replay == eager for new inputs (up to the
    order of the float atomics in the SE-module pooling, which differs from run to run in eager mode too)."""
    import copy
    from common import build_inversion_net
    from invertavatar_b200.graphs import GraphedCall
    net = copy.deepcopy(build_inversion_net(16, 16, 64)).to('cuda')
    net.encoder.eval()                       # (eval-mode BatchNorm: no running-statistics side effects between the two runs)
    x, _, _ = synth.encoder_inputs(2)
    imgs = x['image'].cuda()
    with torch.no_grad():
        gc = GraphedCall(lambda image: net.encode(image), {'image': imgs[:1]})
        for i in (1, 0):
            got = gc(image=imgs[i:i + 1]).clone()
            want = net.encode(imgs[i:i + 1])
            assert float((got - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())), i


@pytest.mark.parametrize('shape', [(2, 3, 16, 16), (3, 3, 64, 128), (1, 32, 8, 12), (2, 3, 6, 6)])
@pytest.mark.parametrize('with_prev', [True, False])
def test_torgb_planar_vec4_matches_element_kernel(monkeypatch, shape, with_prev):
    """The planar ToRGB tail that owns 4 adjacent pixels per thread (torgb_finish_nchw4_kernel: 16-byte plane stores, one
    multimem.st.v4 per 4 values under the fused gather) gives the element-per-thread kernel's values bit for bit; widths that
    are not a multiple of 4 take the element kernel."""
    from invertavatar_b200 import runtime as rt
    B, Cc, H, W = shape
    g = torch.Generator().manual_seed(7)
    raw = (torch.randn(B, H, W, Cc, generator=g) * 200).cuda()        # some values beyond the +-256 clamp
    bias = torch.randn(Cc, generator=g).cuda()
    prev = torch.randn(B, H // 2, W // 2, Cc, generator=g).cuda() if with_prev else None
    monkeypatch.setenv('IA_TORGB_NCHW4', '0')
    want = rt.torgb_finish(raw, bias, 256.0, prev, out_nchw=True).clone()
    monkeypatch.setenv('IA_TORGB_NCHW4', '1')
    got = rt.torgb_finish(raw, bias, 256.0, prev, out_nchw=True)
    assert got.shape == want.shape and torch.equal(got, want)
    nhwc = rt.torgb_finish(raw, bias, 256.0, prev, out_nchw=False)
    assert torch.equal(rt.from_nhwc(nhwc) if hasattr(rt, 'from_nhwc') else nhwc.permute(0, 3, 1, 2), want)


@pytest.mark.parametrize('fmt', ['f16', 'bf16x3', 'both32'])
@pytest.mark.parametrize('B,OH,C,act,noise,demod,clamp', [(2, 128, 64, 'lrelu', True, True, None), (1, 64, 128, 'lrelu', False, True, 256.0),
                                                          (3, 34, 32, 'linear', True, False, None), (1, 70, 256, 'relu', False, False, 1.5)])
def test_fir_epilogue_ring_kernel_bit_identical(monkeypatch, fmt, B, OH, C, act, noise, demod, clamp):
    """The ring kernel (fir_epilogue_x2r_kernel: statically addressed row ring, emission layout as a template parameter) against
    the one-column kernel on the layouts it specialises -- operand 1 alone in single-pass fp16 (backbone up-layers) or bf16 hi/lo
    (super-resolution up-layers) -- and on the generic layout, with and without noise / demodulation / clamp, partial strips
    (OH not a multiple of 32) included: every emitted tensor bit for bit."""
    from invertavatar_b200 import runtime as rt
    g = torch.Generator().manual_seed(B * 1000 + OH + C)
    raw = torch.randn(B, OH + 1, OH + 1, C, generator=g).cuda()
    dcoef = (torch.rand(B, C, generator=g).cuda() + 0.5) if demod else None
    bias = torch.randn(C, generator=g).cuda()
    nz, strength = (torch.randn(OH, OH, generator=g).cuda(), torch.tensor(0.3).cuda()) if noise else (None, None)
    s1 = torch.randn(B, C, generator=g).cuda()
    outs = {}
    for name, env in [('x1', {'IA_FIR_X2': '0', 'IA_FIR_RING': '0'}), ('ring', {'IA_FIR_X2': '1', 'IA_FIR_RING': '1', 'IA_FIR_NOISE_PREFETCH': '1'})]:
        monkeypatch.setenv('IA_FIR_TMA', '0')
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = torch.zeros(B, OH, OH, C, device='cuda') if fmt == 'both32' else None
        e1 = rt.new_split(B, OH, OH, max(64, C), 'cuda', C=C, fmt=rt.FMT_F16X1 if fmt == 'f16' else rt.FMT_BF16X3)
        rt.fir_epilogue(raw, rt.fir4x4_gain4('cuda'), out, dcoef, nz, strength, bias, act, 1.3, clamp, e1=(e1, s1))
        torch.cuda.synchronize()
        outs[name] = [t.clone() for t in ([out] if out is not None else []) + [e1.hi] + ([e1.lo] if e1.lo is not None else [])]
    for a, b in zip(outs['x1'], outs['ring']):
        assert torch.equal(a, b)
