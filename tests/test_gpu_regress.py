"""Regressions found on the GPU."""
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
