"""Host-side logic that needs no GPU: operand layouts, fusion predicates, stream/bench plumbing."""
import json
import os
import subprocess
import sys

import torch

from invertavatar_b200 import runtime as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_padded_split_layout():
    """new_split(pad_row=True): [B][H+1][W][C] buffers with a zero row after every image, exposed as [B,H,W,C] views whose
    image stride is (H+1)*W pixels (ia_emit.e1_img_pix / ia_conv_params.a_img_rows)."""
    B, H, W, C = 3, 4, 8, 64
    sp = rt.new_split(B, H, W, C, 'cpu', pad_row=True)
    assert tuple(sp.hi.shape) == (B, H, W, C) and sp.img_rows == H + 1 and sp.img_pix == (H + 1) * W
    assert sp.hi.stride(0) == (H + 1) * W * C and sp.hi.stride(1) == W * C
    base = sp.hi._base
    assert tuple(base.shape) == (B, H + 1, W, C) and float(base[:, H].float().abs().max()) == 0.0
    dense = rt.new_split(B, H, W, C, 'cpu')
    assert dense.img_pix == 0 and dense.img_rows == H and dense.hi.is_contiguous()
    part = rt.Split(sp.hi[1:], sp.lo[1:], img_rows=sp.img_rows)      # a batch slice keeps the padded stride (grouped prefix)
    assert part.img_pix == sp.img_pix and part.hi.data_ptr() == sp.hi[1].data_ptr()
    padded_c = rt.new_split(B, H, W, C, 'cpu', C=40, pad_row=True)     # channel padding: zero-filled as a whole
    assert float(padded_c.hi._base.float().abs().max()) == 0.0


def test_fusion_predicates(monkeypatch):
    old = rt.get_conv_impl()
    try:
        rt.set_conv_impl('tc')
        assert rt.can_fuse_torgb(512, 512, 128, 3) and rt.can_fuse_torgb(256, 256, 256, 3)
        assert not rt.can_fuse_torgb(256, 256, 512, 3)        # more than two N tiles would make the sum order-dependent
        assert not rt.can_fuse_torgb(256, 256, 128, 32)       # backbone ToRGB (32 / 96 image channels) stays a 1x1 convolution
        assert not rt.can_fuse_torgb(8, 8, 128, 3)            # below the persistent kernel's minimum image
        assert rt.can_fuse_torgb_tail(64, 64, 32) and not rt.can_fuse_torgb_tail(8, 8, 32) and not rt.can_fuse_torgb_tail(64, 64, 3)
        assert rt.pad_row_wanted(32, 32) and not rt.pad_row_wanted(16, 16)
        monkeypatch.setenv('IA_FUSE_TORGB', '0')
        monkeypatch.setenv('IA_FUSE_TORGB_TAIL', '0')
        monkeypatch.setenv('IA_CONV_CAT_ROWS', '0')
        assert not rt.can_fuse_torgb(512, 512, 128, 3) and not rt.can_fuse_torgb_tail(64, 64, 32) and not rt.pad_row_wanted(32, 32)
        rt.set_conv_impl('simt')
        monkeypatch.delenv('IA_FUSE_TORGB'); monkeypatch.delenv('IA_FUSE_TORGB_TAIL'); monkeypatch.delenv('IA_CONV_CAT_ROWS')
        assert not rt.can_fuse_torgb(512, 512, 128, 3) and not rt.can_fuse_torgb_tail(64, 64, 32) and not rt.pad_row_wanted(32, 32)
    finally:
        rt.set_conv_impl(old)


def test_reference_arm_other_ranks_exit_silently():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output (tier contract)."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_bench_stdout_is_one_json_line():
    """Whatever libraries print on fd 1 goes to stderr; stdout carries exactly the JSON line (reference arm, bounded sample)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['cpu_baseline']['kind'] == 'port' and d['e2e']['h2d_bytes_per_step'] == 0


def test_persistent_pickle_loads_in_a_fresh_process(tmp_path):
    """A pickle of an engine @persistent_class object loads in a process that has not imported the defining module
    (class found by its recorded module name), and still loads when the module text differs from the pickled one (rebuilt
    from the pickled source, whose relative imports resolve against the package)."""
    import pickle
    from invertavatar_b200 import persistence
    from invertavatar_b200.stylegan2 import FullyConnectedLayer
    torch.manual_seed(0)
    fc = FullyConnectedLayer(4, 3, lr_multiplier=0.5)
    path = tmp_path / 'fc.pkl'
    with open(path, 'wb') as f:
        pickle.dump(fc, f)
    # same record, but with a module text that no importable module has (as after an upgrade of the package) and -- second
    # variant -- without the module_name field (a pickle written by the previous version of the engine)
    fields = fc.__reduce__()
    meta = dict(fields[1][0])
    meta['module_src'] = meta['module_src'] + '\n# edited after the pickle was written\n'
    path2 = tmp_path / 'fc_changed.pkl'
    with open(path2, 'wb') as f:
        pickle.dump((meta, {k: v for k, v in meta.items() if k != 'module_name'}), f)
    code = '''
import pickle, sys
sys.path.insert(0, %r)
import torch
assert 'invertavatar_b200.stylegan2' not in sys.modules
with open(%r, 'rb') as f:
    fc = pickle.load(f)
assert type(fc).__name__ == 'FullyConnectedLayer' and tuple(fc.weight.shape) == (3, 4), type(fc)
assert fc.init_args == (4, 3) and fc.init_kwargs['lr_multiplier'] == 0.5
from invertavatar_b200 import persistence
with open(%r, 'rb') as f:
    metas = pickle.load(f)
for meta in metas:
    obj = persistence._reconstruct_persistent_obj(meta)
    assert type(obj).__name__ == 'FullyConnectedLayer' and type(obj).__module__.startswith('_imported_module_'), type(obj).__module__
    assert torch.equal(obj.weight, fc.weight)
print('ok')
''' % (ROOT, str(path), str(path2))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stdout + r.stderr


def test_checkpoint_hash_travels_to_the_pack_cache_key(tmp_path, monkeypatch):
    """legacy.load_network_pkl records the sha256 of the pickle bytes on the loaded networks, copy_params_and_buffers(require_all)
    carries it to the rebuilt module, and the pack-cache file name depends on it, on the library's source hash and on the
    precision policy (one-time weight packing cached by checkpoint hash)."""
    import hashlib
    import pickle
    from invertavatar_b200 import glue
    from invertavatar_b200.stylegan2 import FullyConnectedLayer
    torch.manual_seed(0)
    G = FullyConnectedLayer(4, 3)
    path = tmp_path / 'net.pkl'
    with open(path, 'wb') as f:
        pickle.dump({'G': G, 'G_ema': G}, f)
    with open(path, 'rb') as f:
        data = glue.load_network_pkl(f)
    digest = hashlib.sha256(open(path, 'rb').read()).hexdigest()
    assert data['G_ema'].__dict__['_ia_source_hash'] == digest + ':G_ema'
    G2 = FullyConnectedLayer(4, 3)
    glue.copy_params_and_buffers(data['G_ema'], G2, require_all=True)
    assert G2.__dict__['_ia_source_hash'] == digest + ':G_ema' and torch.equal(G2.weight, G.weight)
    a = rt.pack_cache_path(digest + ':G_ema', cache_dir=str(tmp_path))
    b = rt.pack_cache_path(digest + ':G', cache_dir=str(tmp_path))
    monkeypatch.setenv('IA_CONV_PRECISION', 'bf16x3')
    c = rt.pack_cache_path(digest + ':G_ema', cache_dir=str(tmp_path))
    assert a != b and a != c and a.startswith(str(tmp_path)) and a.endswith('.pt')


def test_capture_scope_keys_engine_state_per_graph():
    """Persistent engine state (split-K scratch, SE / BatchNorm sums, StylePlan buffers) is keyed by the capture scope that is active
    when it is requested: scope 0 outside graphs.GraphedCall / GraphedSynthesis, a fresh scope per captured graph, restored on exit
    (nested scopes included) -- so two graphs never share state, and eager calls never share a graph's."""
    import torch
    from invertavatar_b200 import runtime as rt
    dev = torch.device('cuda', 0)
    k0 = rt._scratch_key(dev, 5)
    assert k0 == (0, 5, 0)
    with rt.capture_scope() as s1:
        k1 = rt._scratch_key(dev, 5)
        with rt.capture_scope() as s2:
            k2 = rt._scratch_key(dev, 5)
        assert rt._scratch_key(dev, 5) == k1
    assert rt._scratch_key(dev, 5) == k0
    assert s1 != s2 and len({k0, k1, k2}) == 3 and k1[2] == s1 and k2[2] == s2
    with rt.capture_scope() as s3:
        assert s3 not in (s1, s2)
    plan = rt.StylePlan.__new__(rt.StylePlan)
    assert plan._key(8) == (8, 0)
    with rt.capture_scope() as s4:
        assert plan._key(8) == (8, s4)


def test_cell_merged_rasterizer_algebra():
    """The identity raster_hpass_merge_kernel / raster_fused_kernel rest on, restated with torch on the CPU: an output pixel of
    aa_resize(grid_sample(tex, uv)) is linear in the texels, so summing the weight products ay*ax*w_corner of all samples that fall
    into one texel cell and gathering the cell's four texels once gives the per-sample sum (triplane_v20.py:331-337) up to fp32
    reassociation."""
    import numpy as np
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    Ht = Wt = 8
    Cc, U, r = 6, 32, 8                                   # 32^2 samples -> 8^2 outputs (scale 4: 8 x 8 samples per output window)
    tex = torch.randn(1, Cc, Ht, Wt, generator=g, dtype=torch.float64)
    lin = (torch.arange(U, dtype=torch.float64) + 0.5) / U * 2 - 1
    yy, xx = torch.meshgrid(lin, lin, indexing='ij')
    uv = torch.stack([xx + 0.05 * torch.sin(3 * yy), yy * 1.1 + 0.05 * torch.cos(2 * xx)], dim=-1).unsqueeze(0)   # leaves the texture at the rim
    want = F.interpolate(F.grid_sample(tex, uv, mode='bilinear', padding_mode='zeros', align_corners=False), size=(r, r), mode='bilinear',
                         antialias=True)[0]
    # separable antialias taps: scale 4, support 4, triangle weights normalised per output
    def taps(i):
        scale = U / r
        center = scale * (i + 0.5)
        lo, hi = max(0, int(center - scale + 0.5)), min(U, int(center + scale + 0.5))
        w = torch.tensor([max(0.0, 1 - abs((j - center + 0.5) / scale)) for j in range(lo, hi)], dtype=torch.float64)
        return lo, w / w.sum()
    got = torch.zeros(Cc, r, r, dtype=torch.float64)
    for oy in range(r):
        ys, wy = taps(oy)
        for ox in range(r):
            xs, wx = taps(ox)
            cells = {}                                    # (y0, x0) -> four summed corner weights (zero-padding corners carry none)
            for sy, ay in enumerate(wy):
                for sx, ax in enumerate(wx):
                    gx, gy = float(uv[0, ys + sy, xs + sx, 0]), float(uv[0, ys + sy, xs + sx, 1])
                    ix, iy = ((gx + 1) * Wt - 1) / 2, ((gy + 1) * Ht - 1) / 2
                    x0, y0 = int(np.floor(ix)), int(np.floor(iy))
                    w4 = [(x0 + 1 - ix) * (y0 + 1 - iy), (ix - x0) * (y0 + 1 - iy), (x0 + 1 - ix) * (iy - y0), (ix - x0) * (iy - y0)]
                    acc = cells.setdefault((y0, x0), [0.0, 0.0, 0.0, 0.0])
                    for k in range(4):
                        acc[k] += float(ay * ax) * w4[k]
            assert len(cells) < len(wy) * len(wx) / 3     # the window really collapses onto a handful of cells
            for (y0, x0), w4 in cells.items():
                for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                    yq, xq = y0 + dy, x0 + dx
                    if 0 <= yq < Ht and 0 <= xq < Wt:
                        got[:, oy, ox] += w4[k] * tex[0, :, yq, xq]
    assert float((got - want).abs().max()) < 1e-12
