"""Diagnostic: does one generator step leave reference cycles that hold CUDA tensors (freed only by the cyclic GC)?"""
import gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from invertavatar_b200 import synth
from invertavatar_b200.triplane import TriPlaneGenerator
torch.manual_seed(0)
G = TriPlaneGenerator(**synth.generator_kwargs(48, 48)).eval().requires_grad_(False).cuda()
B = 8
z, cond, c, uv = synth.latents(B).cuda(), synth.frontal_camera(B).cuda(), synth.cameras(B).cuda(), synth.uvcoords_image(B).cuda()
def step():
    ws = G.mapping(z, cond, truncation_psi=0.7, truncation_cutoff=14)
    return G.synthesis(ws, c, {'uvcoords_image': uv}, neural_rendering_resolution=128, noise_mode='const', evaluation=True)['image']
with torch.no_grad():
    for _ in range(3): step()
    torch.cuda.synchronize(); gc.collect()
    gc.disable()
    m0 = torch.cuda.memory_allocated()
    for i in range(4):
        step(); torch.cuda.synchronize()
        print('step', i, 'allocated MB', (torch.cuda.memory_allocated() - m0) / 1e6, 'reserved MB', torch.cuda.memory_reserved() / 1e6)
    gc.set_debug(gc.DEBUG_SAVEALL)
    n = gc.collect()
    held = [o for o in gc.garbage if isinstance(o, torch.Tensor)]
    print('unreachable objects', n, 'tensors among them', len(held), 'bytes', sum(t.numel() * t.element_size() for t in held if t.is_cuda) / 1e6, 'MB')
    import collections
    print(collections.Counter(type(o).__name__ for o in gc.garbage).most_common(12))
