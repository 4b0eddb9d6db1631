"""Per-level timing of the rasterizer variants (two-pass per-sample / row-merged, one-launch fused) at the headline shapes:
python tools/prof_raster.py  (on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import runtime as rt, synth

B = 8
uv = synth.uvcoords_image(B).cuda().contiguous()
variants = [('per_sample', {'IA_RASTER_MERGE': '0', 'IA_RASTER_FUSED': '0'}), ('row_merge', {'IA_RASTER_MERGE': '1', 'IA_RASTER_FUSED': '0'}),
            ('fused', {'IA_RASTER_MERGE': '1', 'IA_RASTER_FUSED': '1'})]
extra = [v for v in sys.argv[1:]]
for Cc, res in [(512, 32), (512, 64), (256, 128)]:
    g = torch.Generator().manual_seed(res)
    tex = torch.randn(B, res, res, Cc, generator=g).cuda()
    stat = torch.randn(B, res, res, Cc, generator=g).cuda()
    sb = [round(i * res / 256) for i in (57, 185, 64, 192)]
    alpha = torch.rand(B, res, res, generator=g).cuda()
    line = []
    for name, env in variants:
        os.environ.update(env)
        for e in extra:
            k, v = e.split('=')
            os.environ[k] = v
        for _ in range(3):
            rt.raster_level(tex, uv, stat, tuple(sb), alpha, res)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            rt.raster_level(tex, uv, stat, tuple(sb), alpha, res)
        t1.record()
        torch.cuda.synchronize()
        line.append('%s %.1f us' % (name, t0.elapsed_time(t1) * 1000 / 20))
    print('C=%d r=%d: ' % (Cc, res) + ', '.join(line), flush=True)
