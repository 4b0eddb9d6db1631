"""Whole-frame CUDA graphs for the per-frame drivers of the inference scripts.

``reenact_avatar_next3d.py:214`` and ``eval_seq.py:212`` call ``G.synthesis`` / ``G.synthesis_withTexture`` once per video
frame at batch 1; at that size a frame is ~170 short launches and the host (Python + ctypes) issues them more slowly than the
GPU runs them.  ``GraphedSynthesis`` captures one call into a CUDA graph (every kernel of libinvertavatar_b200.so is launched
on the capturing stream with static buffers; tensor maps are encoded on the host at capture time) and replays it per frame:

    gs = GraphedSynthesis(G, ws, c, uv)                  # or (..., texture_feats=..., static_feats=...) for the eval_seq driver
    img = gs(c_t, uv_t)['image']                         # per frame: two small copies + one graph launch

The numbers are those of the eager call (same kernels, same order).  Inputs keep the shapes they had at capture."""
import torch

from . import runtime as rt


class GraphedSynthesis:
    def __init__(self, G, ws, c, uvcoords_image, texture_feats=None, static_feats=None, neural_rendering_resolution=None,
                 evaluation=True, warmup=3):
        assert ws.is_cuda, 'GraphedSynthesis needs CUDA tensors'
        self.G = G
        self.ws = ws.detach().clone()
        self.c = c.detach().clone()
        self.uv = uvcoords_image.detach().clone().float()
        self.tex = None if texture_feats is None else [t.detach().clone() for t in texture_feats]
        self.sta = None if static_feats is None else [t.detach().clone() for t in static_feats]
        self.res = neural_rendering_resolution
        self.evaluation = evaluation
        with rt.capture_scope():     # persistent engine state built / requested in here is private to this graph (runtime._scratch_key)
            side = torch.cuda.Stream(device=ws.device)
            side.wait_stream(torch.cuda.current_stream(ws.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(warmup):          # builds every cache (packed weights, style plans, tap tables) outside the capture
                    self._call()
            torch.cuda.current_stream(ws.device).wait_stream(side)
            torch.cuda.synchronize(ws.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out = self._call()

    def _call(self):
        mc = {'uvcoords_image': self.uv}
        if self.tex is not None:
            return self.G.synthesis_withTexture(self.ws, self.tex, self.c, mc, static_feats=self.sta, neural_rendering_resolution=self.res,
                                                noise_mode='const', evaluation=self.evaluation)
        return self.G.synthesis(self.ws, self.c, mc, neural_rendering_resolution=self.res, noise_mode='const', evaluation=self.evaluation)

    def __call__(self, c=None, uvcoords_image=None, ws=None):
        """Replay with new camera / mesh condition / latent (any of them may be omitted); returns the dict of the captured call
        (static output tensors: copy what must outlive the next replay)."""
        if ws is not None:
            self.ws.copy_(ws, non_blocking=True)
        if c is not None:
            self.c.copy_(c, non_blocking=True)
        if uvcoords_image is not None:
            self.uv.copy_(uvcoords_image, non_blocking=True)
        self.graph.replay()
        return self.out


class GraphedCall:
    """A whole call sequence captured into ONE CUDA graph:

        gc = GraphedCall(fn, inputs)        # inputs: dict name -> CUDA tensor (shapes are frozen); fn(**inputs) -> tensor / tuple / dict of tensors
        out = gc(image=new_image, ...)      # copy the given inputs into the static buffers, replay, return the (static) outputs

    Meant for launch-bound sequences of this library -- e.g. one identity of eval_seq.py:164-212 (e4e encode + the two backbones +
    inversionNet.AR_eval_forward + the per-frame synthesis_withTexture calls) is ~2 400 launches whose host cost (~25 us each
    through Python + ctypes) exceeds their device time.  Everything the library does is capture-safe: kernels go to the
    capturing stream(s), side streams fork and join with events, tensor maps are encoded on the host at capture time, scratch
    comes from torch's (graph-private) caching allocator.  Random draws made with torch inside ``fn`` use the graph-safe
    Philox state of torch.cuda.graph (a new draw on every replay).  The numbers are those of the eager call."""

    def __init__(self, fn, inputs, warmup=2):
        dev = next(iter(inputs.values())).device
        assert dev.type == 'cuda', 'GraphedCall needs CUDA tensors'
        self.fn = fn
        self.inputs = {k: v.detach().clone() for k, v in inputs.items()}
        with rt.capture_scope():     # persistent engine state built / requested in here is private to this graph (runtime._scratch_key)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(warmup):          # builds every cache (packed weights, style plans, tap tables, split-K scratch) outside the capture
                    fn(**self.inputs)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out = fn(**self.inputs)

    def __call__(self, **new_inputs):
        for k, v in new_inputs.items():
            self.inputs[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
