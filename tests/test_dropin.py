"""The reference's module paths resolve onto this package through invertavatar_b200/dropin (CPU: import
surface, signatures, the reload idiom of reenact_avatar_next3d.py:156-162 / eval_seq.py:89-97)."""
import importlib
import inspect
import subprocess
import sys

import pytest
import torch

from common import build_generator

PATHS = {
    'training_avatar_texture.triplane_v20': ['TriPlaneGenerator'],
    'training_avatar_texture.networks_stylegan2_new': ['Generator', 'modulated_conv2d', 'SynthesisNetwork'],
    'training_avatar_texture.superresolution': ['SuperresolutionHybrid8XDC', 'SuperresolutionHybrid8X'],
    'training_avatar_texture.volumetric_rendering.renderer': ['ImportanceRenderer', 'ImportanceRenderer_bsMotion', 'fill_mouth',
                                                              'sample_from_planes', 'generate_planes'],
    'training_avatar_texture.volumetric_rendering.ray_sampler': ['RaySampler', 'RaySampler_zxc'],
    'training_avatar_texture.volumetric_rendering.ray_marcher': ['MipRayMarcher2'],
    'training_avatar_texture.camera_utils': ['LookAtPoseSampler', 'FOV_to_intrinsics'],
    'training.networks_stylegan2': ['FullyConnectedLayer', 'SynthesisLayer', 'ToRGBLayer', 'SynthesisBlock', 'modulated_conv2d'],
    'torch_utils.ops.bias_act': ['bias_act', 'activation_funcs'],
    'torch_utils.ops.upfirdn2d': ['upfirdn2d', 'setup_filter', 'filter2d', 'upsample2d', 'downsample2d'],
    'torch_utils.ops.filtered_lrelu': ['filtered_lrelu'],
    'torch_utils.ops.conv2d_resample': ['conv2d_resample'],
    'torch_utils.ops.conv2d_gradfix': ['conv2d', 'conv_transpose2d', 'enabled', 'no_weight_gradients'],
    'torch_utils.ops.fma': ['fma'],
    'torch_utils.ops.grid_sample_gradfix': ['grid_sample', 'enabled'],
    'torch_utils.misc': ['copy_params_and_buffers', 'assert_shape', 'profiled_function'],
    'torch_utils.persistence': ['persistent_class'],
    'torch_utils.custom_ops': ['get_plugin'],
    'encoder_inversion.models.uvnet': ['inversionNet'],
    'encoder_inversion.models.unet_encoders': ['ConvGRU', 'TriPlanefeat_Encoder', 'TriPlaneSFTfeat_Encoder'],
    'encoder_inversion.models.e4e': ['Encoder4Editing'],
    'encoder_inversion.models.uvnet_new': ['inversionNet', 'improved_os_unet_encoder'],
    'encoder_inversion.models.unet_transformer': ['UpLayer', 'TriPlanefeat_SegformerDecoder', 'TriPlaneSFTfeat_SegformerDecoder'],
    'encoder_inversion.models.mmseg.mix_transformer': ['MixVisionTransformer', 'Block', 'Attention', 'Mlp', 'OverlapPatchEmbed', 'transformer_block', 'MLP'],
    'dnnlib.util': ['EasyDict', 'construct_class_by_name', 'open_url'],
    'legacy': ['load_network_pkl'],
}


def test_module_paths_resolve_in_a_fresh_interpreter():
    """Fresh process with only the drop-in tree in front of sys.path: every path of SURVEY 8(b) imports and exports its names."""
    code = 'import invertavatar_b200.dropin as d; d.install()\nimport importlib\n'
    for mod, names in PATHS.items():
        code += f'm = importlib.import_module({mod!r})\n'
        for n in names:
            code += f'assert hasattr(m, {n!r}), ({mod!r}, {n!r})\n'
    code += 'import training_avatar_texture.triplane_v20 as t; assert t.__file__.find("dropin") > 0\nprint("ok")\n'
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=str(__import__('pathlib').Path(__file__).resolve().parents[1]))
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]


def test_signatures_follow_the_reference():
    from invertavatar_b200 import ops
    from invertavatar_b200.triplane import TriPlaneGenerator
    assert list(inspect.signature(ops.bias_act).parameters) == ['x', 'b', 'dim', 'act', 'alpha', 'gain', 'clamp', 'impl']
    assert list(inspect.signature(ops.upfirdn2d).parameters) == ['x', 'f', 'up', 'down', 'padding', 'flip_filter', 'gain', 'impl']
    assert list(inspect.signature(ops.filtered_lrelu).parameters) == ['x', 'fu', 'fd', 'b', 'up', 'down', 'padding', 'gain', 'slope', 'clamp',
                                                                      'flip_filter', 'impl']
    assert list(inspect.signature(ops.conv2d_resample).parameters) == ['x', 'w', 'f', 'up', 'down', 'padding', 'groups', 'flip_weight', 'flip_filter']
    syn = list(inspect.signature(TriPlaneGenerator.synthesis).parameters)
    assert syn[:10] == ['self', 'ws', 'c', 'mesh_condition', 'neural_rendering_resolution', 'update_emas', 'cache_backbone',
                        'use_cached_backbone', 'return_featmap', 'evaluation']
    assert ops.activation_funcs['lrelu'].def_alpha == 0.2 and abs(ops.activation_funcs['lrelu'].def_gain - 2 ** 0.5) < 1e-12
    assert ops.activation_funcs['sigmoid'].cuda_idx == 5


def test_reload_idiom_and_cpu_inputs_fail_loudly():
    """TriPlaneGenerator(*G.init_args, **G.init_kwargs) + copy_params_and_buffers(require_all=True); CPU tensors raise."""
    from invertavatar_b200 import glue, ops
    from invertavatar_b200.triplane import TriPlaneGenerator
    G = build_generator(16, 16)
    G2 = TriPlaneGenerator(*G.init_args, **G.init_kwargs).eval().requires_grad_(False)
    glue.copy_params_and_buffers(G, G2, require_all=True)
    for (k1, v1), (k2, v2) in zip(G.state_dict().items(), G2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert len(G.state_dict()) == 444
    bad = torch.nn.Linear(2, 2)
    with pytest.raises(AssertionError):
        glue.copy_params_and_buffers(bad, G2, require_all=True)
    with pytest.raises(RuntimeError):
        ops.bias_act(torch.zeros(1, 2, 3, 3), torch.zeros(2))
    with pytest.raises(RuntimeError):
        ops.upfirdn2d(torch.zeros(1, 2, 4, 4), ops.setup_filter([1, 3, 3, 1]), up=2)


def test_engine_caches_do_not_block_copy_or_pickle():
    """copy.deepcopy / pickle of a module that carries engine-side caches (ctypes tables, packed weights) drops the caches."""
    import copy
    import pickle
    from invertavatar_b200 import runtime as rt
    lin = torch.nn.Linear(2, 2)
    plan = rt.StylePlan([], torch.device('cpu'))
    import ctypes
    plan.host = (ctypes.c_void_p * 2)()        # what a used plan holds
    lin.__dict__['_ia_plan'] = (('k',), plan)
    lin2 = copy.deepcopy(lin)
    assert lin2.__dict__['_ia_plan'][1] is None
    lin3 = pickle.loads(pickle.dumps(lin))
    assert lin3.__dict__['_ia_plan'][1] is None and torch.equal(lin3.weight, lin.weight)
