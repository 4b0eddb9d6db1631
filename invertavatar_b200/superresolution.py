"""Super-resolution modules (reference training_avatar_texture/superresolution.py), same constructor signatures and
state-dict names, running on the B200 engine.  The reference computes these blocks in fp16 on CUDA and fp32 on CPU with
clamp 256 either way (superresolution.py:270-276); the engine computes them like every other block:
3-term split bf16 tensor-core convolution with fp32 accumulation, i.e. at the precision of the reference's fp32 path."""
import torch

from . import persistence
from . import runtime as rt
from .stylegan2 import SynthesisBlock, _plan_for


class _SuperresolutionBase(torch.nn.Module):
    input_resolution = 128

    def _prepare(self, rgb, x):
        """NCHW (any strides) -> NHWC, antialiased resize to the module's input resolution when needed
        (superresolution.py:281-285)."""
        xn, rn = rt.to_nhwc(x), rt.to_nhwc(rgb)
        if xn.shape[2] != self.input_resolution:
            if not self.sr_antialias:
                raise NotImplementedError('sr_antialias=False resize is not implemented (the generator sets it True)')
            r = self.input_resolution
            xn, rn = rt.resize_aa(xn, r, r), rt.resize_aa(rn, r, r)
        return rn, xn

    def forward(self, rgb, x, ws, **block_kwargs):
        ws = ws[:, -1:, :].repeat(1, 3, 1)
        noise_mode = block_kwargs.get('noise_mode', 'random')
        rgb, x = self._prepare(rgb, x)
        B = x.shape[0]
        b0, b1 = self.block0, self.block1
        entries = [layer.style_entry(j) for j, (_, layer) in enumerate(b0.layers())] + \
                  [layer.style_entry(j) for j, (_, layer) in enumerate(b1.layers())]
        styles, dcoefs = _plan_for(self, entries).run(ws.to(torch.float32))
        n0 = len(b0.layers())
        a0 = rt.modsplit_split(x, styles[0], C_pad=b0.conv0.pack().Cin_pad, pad_row=rt.pad_row_wanted(x.shape[1], x.shape[2]))
        _, rgb, a = b0.run_chain(a0, rgb, styles[:n0], dcoefs[:n0], B, noise_mode=noise_mode, next_conv=b1.conv0,
                                 next_styles=styles[n0])
        _, rgb, _ = b1.run_chain(a, rgb, styles[n0:], dcoefs[n0:], B, noise_mode=noise_mode, img_nchw=True)
        return rgb


def _make(block0_out, block1_out, in_res, res0, res1, img_res):
    def init(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None,
             channel_base=None, channel_max=None, **block_kwargs):
        torch.nn.Module.__init__(self)
        assert img_resolution == img_res
        use_fp16 = sr_num_fp16_res > 0
        self.input_resolution = in_res
        self.sr_antialias = sr_antialias
        self.block0 = SynthesisBlock(channels, block0_out, w_dim=512, resolution=res0, img_channels=3, is_last=False,
                                     use_fp16=use_fp16, conv_clamp=(256 if use_fp16 else None), **block_kwargs)
        self.block1 = SynthesisBlock(block0_out, block1_out, w_dim=512, resolution=res1, img_channels=3, is_last=True,
                                     use_fp16=use_fp16, conv_clamp=(256 if use_fp16 else None), **block_kwargs)
    return init


@persistence.persistent_class
class SuperresolutionHybrid8XDC(_SuperresolutionBase):
    """128^2 x 32ch -> 512^2 x 3: block0 32->256 @256^2, block1 256->128 @512^2 (superresolution.py:263-289)."""
    __init__ = _make(256, 128, 128, 256, 512, 512)


@persistence.persistent_class
class SuperresolutionHybrid8X(_SuperresolutionBase):
    """superresolution.py:28-58 (128/64 channels; carries an extra resample_filter buffer)."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None,
                 channel_base=None, channel_max=None, **block_kwargs):
        _make(128, 64, 128, 256, 512, 512)(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, **block_kwargs)
        from .stylegan2 import setup_filter
        self.register_buffer('resample_filter', setup_filter([1, 3, 3, 1]))
