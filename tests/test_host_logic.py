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
