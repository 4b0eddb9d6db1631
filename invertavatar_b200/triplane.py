"""Next3D++ tri-plane generator on the B200 engine (reference training_avatar_texture/triplane_v20.py).

Same constructor, attributes, methods, return dictionaries and state-dict (444 tensors) as the reference
``TriPlaneGenerator``; every stage underneath is a kernel of libinvertavatar_b200.so (see DESIGN.md for the stage map).
"""
import importlib

import torch

from . import persistence
from . import runtime as rt
from . import stylegan2 as sg
from . import superresolution as sr_mod
from .rendering import ImportanceRenderer_bsMotion, RaySampler_zxc
from .rendering import fill_mouth as rendering_fill_mouth
from .stylegan2 import FullyConnectedLayer
from .stylegan2 import Generator as StyleGAN2Backbone_cond

BBOX_256 = [57, 185, 64, 192]   # face region of the frontal plane, triplane_v20.py:114

_KNOWN_SR = {
    'training_avatar_texture.superresolution.SuperresolutionHybrid8XDC': sr_mod.SuperresolutionHybrid8XDC,
    'training_avatar_texture.superresolution.SuperresolutionHybrid8X': sr_mod.SuperresolutionHybrid8X,
}


def _grouped_prefix_enabled():
    """IA_GROUPED_PREFIX=0 evaluates the three backbones one after the other (cross-check / profiling)."""
    import os
    return os.environ.get('IA_GROUPED_PREFIX', '1') != '0'


def _prune_dead_enabled():
    import os
    return os.environ.get('IA_PRUNE_DEAD', '1') != '0'


_backbone_streams = None


def set_backbone_streams(enabled):
    """Issue the texture and static backbones on two streams (True, default) or one after the other on the caller's stream
    (False; used when kernels are timed one by one).  None restores the IA_BACKBONE_STREAMS environment default."""
    global _backbone_streams
    _backbone_streams = enabled


def _backbone_streams_enabled():
    if _backbone_streams is not None:
        return bool(_backbone_streams)
    import os
    return os.environ.get('IA_BACKBONE_STREAMS', '1') != '0'


def _construct(class_name, **kwargs):
    """dnnlib.util.construct_class_by_name for the super-resolution module (triplane_v20.py:56-58)."""
    cls = _KNOWN_SR.get(class_name)
    if cls is None:
        mod, name = class_name.rsplit('.', 1)
        cls = getattr(importlib.import_module(mod), name)
    return cls(**kwargs)


class _StageOutputs(dict):
    """Result dictionary of synthesis_withTexture / synthesis_withCondition.  The reference returns the blended tri-planes
    under 'triplane' from these calls (triplane_v20.py:243-244,311); the renderer here consumes them in fp16 and the
    inference scripts never read the entry (eval_seq.py:212 takes ['image']), so the fp32 [B,3,32,256,256] tensor is
    materialised on first access instead of on every frame."""

    def __init__(self, items, make_triplane):
        super().__init__(items)
        self._make_triplane = make_triplane

    def __missing__(self, key):
        if key == 'triplane' and self._make_triplane is not None:
            self['triplane'] = self._make_triplane()
            return dict.__getitem__(self, 'triplane')
        raise KeyError(key)


class OSGDecoder(torch.nn.Module):
    """triplane_v20.py:415-438.  The forward used by the generator is fused into the render kernel; the module holds the
    parameters (decoder.net.{0,2}.{weight,bias}) and offers the standalone forward for API parity."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = torch.nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']),
            torch.nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']))

    def forward(self, sampled_features, ray_directions, sampled_embeddings=None):
        x = sampled_features.mean(1)
        N, M, Cc = x.shape
        x = self.net[0](x.reshape(N * M, Cc))
        x = rt.bias_act(x, act='softplus')
        x = self.net[2](x).reshape(N, M, -1)
        rgb = rt.bias_act(x[..., 1:].contiguous(), act='sigmoid') * (1 + 2 * 0.001) - 0.001
        return {'rgb': rgb, 'sigma': x[..., 0:1]}


@persistence.persistent_class
class TriPlaneGenerator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, topology_path=None, sr_num_fp16_res=0,
                 mapping_kwargs={}, rendering_kwargs={}, sr_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.renderer = ImportanceRenderer_bsMotion()
        self.ray_sampler = RaySampler_zxc()
        self.texture_backbone = StyleGAN2Backbone_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32,
                                                       mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.face_backbone = StyleGAN2Backbone_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32,
                                                    mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.backbone = StyleGAN2Backbone_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                               mapping_ws=self.texture_backbone.num_ws, mapping_kwargs=mapping_kwargs,
                                               **synthesis_kwargs)
        self.superresolution = _construct(rendering_kwargs['superresolution_module'], channels=32, img_resolution=img_resolution,
                                          sr_num_fp16_res=sr_num_fp16_res, sr_antialias=rendering_kwargs['sr_antialias'],
                                          **sr_kwargs)
        self.decoder = OSGDecoder(32, {'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32})
        self.neural_rendering_resolution = 128
        self.rendering_kwargs = rendering_kwargs
        self.fill_mouth = True
        # Per-layer tensor-core precision (rt.layer_fmt): the 3x3 convolutions of the three backbones run as single-pass fp16
        # MMAs (one third of the tensor-core work of the 3-term bf16 split), the super-resolution blocks and every ToRGB keep
        # the 3-term split.  Measured with the CPU oracle on the headline configuration (tools/probe_conv_precision.py,
        # profiles/r2_conv_precision_probe*.json): each backbone layer alone moves the final image by <= 6e-5, all 39 together
        # by 0.9e-4 .. 3.1e-4 max-abs (PSNR 88 .. 98 dB) over 6 frames -- inside the 1e-3 / 50 dB bar with a 3x margin -- whereas
        # ONE super-resolution layer in fp16 costs 1e-3.  IA_CONV_PRECISION=bf16x3 restores the 3-term split everywhere.
        for net in (self.texture_backbone, self.face_backbone, self.backbone):
            for m in net.synthesis.modules():
                if isinstance(m, sg.SynthesisLayer):
                    m.tc_fmt = rt.FMT_F16X1
        # the renderer's decoder MLP as single-pass fp16 mma.sync as well: +1.2e-4 on the final image on its own, 1.2e-4 .. 3.6e-4
        # (PSNR 86 .. 95 dB) together with the backbone layers (profiles/r2_conv_precision_probe_mix_mlp.json)
        self.renderer.mlp_fmt = rt.FMT_F16X1
        # ... and the tri-planes it gathers from are stored as fp16 (9.4e-6 on the final image, profiles/r1_render_precision_probe.json)
        self.renderer.planes_fmt = rt.FMT_F16X1

    def _side_streams(self, device):
        return rt.side_streams(device)

    # ------------------------------------------------------------------------------------------
    def mapping(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        c = c[:, :self.c_dim]
        return self.backbone.mapping(z, c * self.rendering_kwargs.get('c_scale', 0), truncation_psi=truncation_psi,
                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _uv_prep(uvcoords_image, resolutions, want_alpha128=False):
        """Everything the rasterizer / stitcher derives from the mesh condition alone: the flood-filled masks
        (renderer.py:716-741) and the antialias-resized alpha / upper-face alpha of every level (triplane_v20.py:331-337).
        Independent of the backbones, so ``synthesis`` issues it on a side stream while they run."""
        uv = uvcoords_image if uvcoords_image.dtype == torch.float32 else uvcoords_image.float()
        uv = uv.contiguous()
        full_alpha, mouth, upper_alpha = rt.fill_mouth(uv, upper_row0=87)
        alpha4 = uv[..., 2:3]                    # [B,H,W,1] NHWC view with pixel stride 3
        upper4 = upper_alpha.unsqueeze(-1)
        resized = {}
        for res in resolutions:
            if res not in resized:
                resized[res] = (rt.resize_aa(alpha4, res, res).squeeze(-1), rt.resize_aa(upper4, res, res).squeeze(-1))
        prep = {'uv': uv, 'full_alpha': full_alpha, 'mouth': mouth, 'resized': resized}
        if want_alpha128:
            prep['alpha128'] = rt.resize_aa(full_alpha.unsqueeze(-1), 128, 128).squeeze(-1)
        return prep

    @staticmethod
    def _prep_tensors(prep):
        out = [prep['uv'], prep['full_alpha'], prep['mouth']] + [t for pair in prep['resized'].values() for t in pair]
        if 'alpha128' in prep:
            out.append(prep['alpha128'])
        return out

    def _rasterize_nhwc(self, texture_feats, uvcoords_image, static_feats, bbox_256, levels=None, prep=None):
        """Engine rasterizer.  texture_feats/static_feats: lists of NHWC tensors (static entries already reduced to the
        32 plane-0 channels where the reference slices them).  Returns per-level (cond NHWC [B,r,r,C], alpha [B,r,r]),
        full_alpha [B,256,256], mouth [B,256,256].  ``levels`` restricts the work to the entries a consumer reads;
        ``prep`` is the result of ``_uv_prep`` when the caller already has it."""
        idxs = [i for i in range(len(texture_feats)) if levels is None or i in levels]
        if prep is None:
            prep = self._uv_prep(uvcoords_image, [texture_feats[i].shape[1] for i in idxs])
        uv, full_alpha, mouth, resized_alpha = prep['uv'], prep['full_alpha'], prep['mouth'], prep['resized']
        for i in idxs:
            res = texture_feats[i].shape[1]
            if res not in resized_alpha:         # a level the caller's prep did not cover
                resized_alpha[res] = self._uv_prep(uvcoords_image, [res])['resized'][res]
        outs = []
        for idx, tex in enumerate(texture_feats):
            if idx not in idxs:
                outs.append(None)
                continue
            res = tex.shape[1]
            bbox = [round(i * res / 256) for i in bbox_256]
            a_r, ua_r = resized_alpha[res]
            # grid_sample @256^2 -> aa-resize to res -> blend with the resized static crop, fused (no [B,256,256,C] tensor)
            outs.append((rt.raster_level(tex, uv, static_feats[idx], (bbox[0], bbox[1], bbox[2], bbox[3]), a_r, res), ua_r))
        return outs, full_alpha, mouth

    def rasterize(self, texture_feats, uvcoords_image, static_feats, bbox_256):
        """Reference signature (triplane_v20.py:317-339): NCHW lists in, list of [B,C+1,res,res], full_alpha, mouth_masks."""
        tex = [rt.to_nhwc(t) for t in texture_feats]
        sta = [rt.to_nhwc(t) for t in static_feats]
        outs, full_alpha, mouth = self._rasterize_nhwc(tex, uvcoords_image, sta, bbox_256)
        images = [rt.from_nhwc(torch.cat([o[0], o[1].unsqueeze(-1)], dim=-1)) for o in outs]
        return images, full_alpha.unsqueeze(1), mouth.unsqueeze(1)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _static_views(static_feats):
        """NHWC views of the six static features; entries 0 and 5 (the 96-channel images) reduced to plane 0
        (triplane_v20.py:109-112)."""
        sf = [rt.to_nhwc(t) for t in static_feats]
        plane_img = sf[-1]
        views = list(sf)
        if views[0].shape[-1] == 96:
            views[0] = views[0][..., :32]
        views[-1] = plane_img[..., :32] if plane_img.shape[-1] == 96 else plane_img
        return views, plane_img

    def _stitch_render_sr(self, ws, c, mesh_condition, texture_feats, static_feats, neural_rendering_resolution,
                          evaluation, synthesis_kwargs, face_prefix=None, uv_prep=None):
        cam = c[:, -25:]
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        N = ws.shape[0]
        tex = [rt.to_nhwc(t) for t in texture_feats]
        static_views, plane_img = self._static_views(static_feats)
        assert len(tex) >= 4 and len(static_views) >= 4     # (a pruned texture list ends with the last level the rasterizer reads)
        assert plane_img.shape[-1] == 96, 'static backbone must emit 3 x 32 plane channels'

        # UV rasterize: only the four levels the face backbone consumes (cond_list[0..3], networks_stylegan2_new.py:536-540)
        conds, full_alpha, _ = self._rasterize_nhwc(tex, mesh_condition['uvcoords_image'], static_views, BBOX_256,
                                                    levels=(0, 1, 2, 3), prep=uv_prep)
        noise_kwargs = {k: v for k, v in synthesis_kwargs.items() if k in ('noise_mode',)}
        stitch = self.face_backbone.synthesis(ws, cond_list=conds[:4], return_list=False, prefix=face_prefix, **noise_kwargs)   # NCHW view
        stitch = rt.to_nhwc(stitch)                                                                          # [B,256,256,32]

        # stitch into plane 0 of a copy of the static planes (triplane_v20.py:119-128)
        b0, b1, b2, b3 = BBOX_256
        stitch128 = rt.resize_aa(stitch, 128, 128)
        alpha128 = uv_prep['alpha128'] if (uv_prep is not None and 'alpha128' in uv_prep) else \
            rt.resize_aa(full_alpha.unsqueeze(-1), 128, 128).squeeze(-1)
        # one pass: copy of the static planes with the face window of plane 0 blended in, written in the renderer's storage
        # format (fp16 by the generator's policy: halves the gather traffic, 9.4e-6 on the image; fp32 in strict mode or when
        # the caller wants the tri-plane back)
        want_planes32 = bool(synthesis_kwargs.get('_want_triplane', False))
        planes_fp16 = self.renderer._planes_fp16() and not want_planes32
        planes = rt.stitch_planes(plane_img, stitch128, alpha128, (b0, b2), fp16=planes_fp16)

        if evaluation:
            assert synthesis_kwargs.get('noise_mode') == 'const', ('noise_mode' in synthesis_kwargs, synthesis_kwargs.get('noise_mode'))
        res = int(self.neural_rendering_resolution)
        feat, depth, wsum = self.renderer.render_nhwc(planes, self.decoder, cam, res, self.rendering_kwargs, evaluation=evaluation)
        feature_image = rt.from_nhwc(feat)                       # [B,32,res,res] view
        depth_image = depth.reshape(N, 1, res, res)
        rgb_image = feature_image[:, :3]
        sr_kwargs = {k: v for k, v in synthesis_kwargs.items() if k != 'noise_mode' and k not in ('update_emas', '_want_triplane')}
        sr_image = self.superresolution(rgb_image, feature_image, ws, noise_mode=self.rendering_kwargs['superresolution_noise_mode'],
                                        **sr_kwargs)
        out = {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'feature_image': feature_image}
        if planes.dtype == torch.float32:
            out['triplane'] = rt.from_nhwc(planes).reshape(N, 3, 32, planes.shape[1], planes.shape[2])
            return out

        def make_triplane():
            p32 = rt.stitch_planes(plane_img, stitch128, alpha128, (b0, b2), fp16=False)
            return rt.from_nhwc(p32).reshape(N, 3, 32, p32.shape[1], p32.shape[2])
        return _StageOutputs(out, make_triplane)

    def synthesis(self, ws, c, mesh_condition, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, return_featmap=False, evaluation=False, depth_jitter=None, importance_u=None,
                  **synthesis_kwargs):
        """triplane_v20.py:89-150.  ``depth_jitter`` / ``importance_u`` optionally pin the renderer's random draws."""
        if depth_jitter is not None:
            self.renderer.depth_jitter = depth_jitter
        if importance_u is not None:
            self.renderer.importance_u = importance_u
        noise_kwargs = {k: v for k, v in synthesis_kwargs.items() if k in ('noise_mode',)}
        # The blocks up to 32^2 of the three backbones do not depend on each other (the face backbone receives its
        # conditions after its own 32^2 block) and are latency-bound: evaluate them as one grouped batch.
        nets = [self.texture_backbone.synthesis, self.backbone.synthesis, self.face_backbone.synthesis]
        # Dead-code elimination in the texture backbone: of its return_list [img32, x32, x64, x128, x256, img256] the rasterizer
        # reads the first four (cond_list of the face backbone, networks_stylegan2_new.py:536-540); the 256^2 block and the
        # skip-connection images after img32 feed nothing unless the caller asked for the feature maps (return_featmap).  The
        # reference evaluates them and drops them; here they are not launched (IA_PRUNE_DEAD=0 evaluates them anyway).
        tex_kwargs = dict(noise_kwargs)
        if not return_featmap and _prune_dead_enabled():
            tex_kwargs['prune_after'] = 128
        pre = [None, None, None]
        if _grouped_prefix_enabled() and sg.can_group_prefix(nets):
            pre = sg.synthesis_prefix_grouped(nets, ws, noise_mode=noise_kwargs.get('noise_mode', 'random'))
        if _backbone_streams_enabled() and ws.is_cuda:
            # The texture and static backbones are independent: issue them on two streams so that the HBM-bound kernels of one
            # (FIR epilogues, ToRGB tails, operand preparation: no shared memory, few registers) run on the SMs next to the
            # tensor-core convolutions of the other; join before the rasterizer consumes both.
            cur = torch.cuda.current_stream(ws.device)
            s_a, s_b, s_c = rt.side_streams(ws.device, 3)
            s_a.wait_stream(cur)
            s_b.wait_stream(cur)
            # third stream: what the rasterizer / stitcher derives from the mesh condition alone (flood fill, alpha resizes) --
            # small latency-bound launches that fit next to the convolutions
            s_c.wait_stream(cur)
            with torch.cuda.stream(s_c):
                uv_prep = self._uv_prep(mesh_condition['uvcoords_image'], (32, 64, 128), want_alpha128=True)   # levels 0..3 of the texture list
            with torch.cuda.stream(s_a):
                texture_feats = self.texture_backbone.synthesis(ws, cond_list=None, return_list=True, prefix=pre[0], **tex_kwargs)
            with torch.cuda.stream(s_b):
                static_feats = self.backbone.synthesis(ws, cond_list=None, return_list=True, prefix=pre[1], **noise_kwargs)
            cur.wait_stream(s_a)
            cur.wait_stream(s_b)
            cur.wait_stream(s_c)
            for t in list(texture_feats) + list(static_feats) + self._prep_tensors(uv_prep):
                t.record_stream(cur)
        else:
            uv_prep = None
            texture_feats = self.texture_backbone.synthesis(ws, cond_list=None, return_list=True, prefix=pre[0], **tex_kwargs)
            static_feats = self.backbone.synthesis(ws, cond_list=None, return_list=True, prefix=pre[1], **noise_kwargs)
        if return_featmap:
            synthesis_kwargs = dict(synthesis_kwargs, _want_triplane=True)      # the caller reads out['triplane']: fp32 planes
        out = self._stitch_render_sr(ws, c, mesh_condition, texture_feats, static_feats, neural_rendering_resolution,
                                     evaluation, synthesis_kwargs, face_prefix=pre[2], uv_prep=uv_prep)
        if return_featmap:
            out['texture'] = texture_feats
            return out
        return {'image': out['image'], 'image_raw': out['image_raw'], 'image_depth': out['image_depth']}

    def synthesis_withTexture(self, ws, texture_feats, c, mesh_condition, static_feats=None, neural_rendering_resolution=None,
                              update_emas=False, cache_backbone=False, use_cached_backbone=False, evaluation=False,
                              depth_jitter=None, importance_u=None, **synthesis_kwargs):
        """triplane_v20.py:152-244: per-frame driver of eval_seq.py (texture/static features precomputed)."""
        if depth_jitter is not None:
            self.renderer.depth_jitter = depth_jitter
        if importance_u is not None:
            self.renderer.importance_u = importance_u
        if static_feats is None:
            noise_kwargs = {k: v for k, v in synthesis_kwargs.items() if k in ('noise_mode',)}
            static_feats = self.backbone.synthesis(ws, cond_list=None, return_list=True, **noise_kwargs)
        return self._stitch_render_sr(ws, c, mesh_condition, texture_feats, static_feats, neural_rendering_resolution,
                                      evaluation, synthesis_kwargs)

    def synthesis_withCondition(self, ws, c, mesh_condition, gt_texture_feats=None, gt_static_feats=None,
                                texture_feats_conditions=None, static_feats_conditions=None, neural_rendering_resolution=None,
                                update_emas=False, cache_backbone=False, use_cached_backbone=False, only_image=False,
                                return_feats=False, **synthesis_kwargs):
        """triplane_v20.py:246-315 (called from training code only; kept for signature parity)."""
        noise_kwargs = {k: v for k, v in synthesis_kwargs.items() if k in ('noise_mode',)}
        texture_feats = gt_texture_feats if gt_texture_feats is not None else self.texture_backbone.synthesis(
            ws, cond_list=None, return_list=True, feat_conditions=texture_feats_conditions, **noise_kwargs)
        static_feats = gt_static_feats if gt_static_feats is not None else self.backbone.synthesis(
            ws, cond_list=None, return_list=True, feat_conditions=static_feats_conditions, **noise_kwargs)
        evaluation = synthesis_kwargs.get('noise_mode') == 'const'
        out = self._stitch_render_sr(ws, c, mesh_condition, texture_feats, static_feats, neural_rendering_resolution,
                                     evaluation, synthesis_kwargs)
        if only_image:
            return {'image': out['image']}
        if return_feats:
            out['static'] = static_feats
            out['texture'] = texture_feats
        return out

    def _blended_planes(self, ws, mesh_condition, synthesis_kwargs):
        """The tri-planes a frame is rendered from (backbones -> rasterize -> face backbone -> stitch), fp32 [B,3,32,256,256]."""
        noise_kwargs = {k: v for k, v in synthesis_kwargs.items() if k in ('noise_mode',)}
        texture_feats = self.texture_backbone.synthesis(ws, cond_list=None, return_list=True, **noise_kwargs)
        static_feats = self.backbone.synthesis(ws, cond_list=None, return_list=True, **noise_kwargs)
        tex = [rt.to_nhwc(t) for t in texture_feats]
        static_views, plane_img = self._static_views(static_feats)
        conds, full_alpha, _ = self._rasterize_nhwc(tex, mesh_condition['uvcoords_image'], static_views, BBOX_256, levels=(0, 1, 2, 3))
        stitch = rt.to_nhwc(self.face_backbone.synthesis(ws, cond_list=conds[:4], return_list=False, **noise_kwargs))
        b0, b1, b2, b3 = BBOX_256
        planes = rt.stitch_planes(plane_img, rt.resize_aa(stitch, 128, 128), rt.resize_aa(full_alpha.unsqueeze(-1), 128, 128).squeeze(-1),
                                  (b0, b2), fp16=False)
        N = ws.shape[0]
        return rt.from_nhwc(planes).reshape(N, 3, 32, planes.shape[1], planes.shape[2])

    def sample(self, coordinates, directions, z, c, mesh_condition, truncation_psi=1, truncation_cutoff=None, update_emas=False,
               **synthesis_kwargs):
        """triplane_v20.py:341-371: decoder outputs {'rgb', 'sigma'} at arbitrary 3-D coordinates [B,M,3] (shape extraction)."""
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.sample_mixed(coordinates, directions, ws, mesh_condition, update_emas=update_emas, **synthesis_kwargs)

    def sample_mixed(self, coordinates, directions, ws, mesh_condition, truncation_psi=1, truncation_cutoff=None, update_emas=False,
                     **synthesis_kwargs):
        """triplane_v20.py:373-402: as ``sample`` but from W+ latents."""
        planes = self._blended_planes(ws, mesh_condition, synthesis_kwargs)
        return self.renderer.run_model(planes, self.decoder, coordinates, directions, self.rendering_kwargs)

    def visualize_mesh_condition(self, mesh_condition, to_imgs=False):
        """triplane_v20.py:71-87: the UV-coordinate image [B,3,H,W] with everything outside the mouth-filled mask set to -1
        (uint8 PIL images with to_imgs=True)."""
        uv = mesh_condition['uvcoords_image'].clone().permute(0, 3, 1, 2)
        full_alpha, _ = rendering_fill_mouth(uv[:, 2:].clone(), blur_mouth_edge=False)
        if not to_imgs:
            return uv
        uv[full_alpha.expand(-1, 3, -1, -1) == 0] = -1
        u8 = ((uv + 1) * 127.5).to(dtype=torch.uint8).cpu()
        from PIL import Image
        return [Image.fromarray(img.permute(1, 2, 0).numpy()) for img in u8]

    def forward(self, z, c, v, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None, update_emas=False,
                cache_backbone=False, use_cached_backbone=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, c, v, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution,
                              cache_backbone=cache_backbone, use_cached_backbone=use_cached_backbone, **synthesis_kwargs)
