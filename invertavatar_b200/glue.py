"""Runtime glue the inference scripts import next to the hot path, API-compatible with the reference's
``torch_utils/misc.py``, ``dnnlib/util.py``, ``legacy.py`` and ``training_avatar_texture/camera_utils.py`` (host-side
bookkeeping only; written against their behaviour)."""
import contextlib
import importlib
import io
import math
import os
import pickle
import re

import numpy as np
import torch

from . import persistence
from . import synth
from .persistence import EasyDict  # noqa: F401


# ---- torch_utils/misc.py ---------------------------------------------------------------------------------------------
_constant_cache = {}


def constant(value, shape=None, dtype=None, device=None, memory_format=None):
    """Cached constant tensor (misc.py:24-46)."""
    value = np.asarray(value)
    shape = tuple(shape) if shape is not None else None
    dtype = dtype or torch.get_default_dtype()
    device = torch.device(device or 'cpu')
    memory_format = memory_format or torch.contiguous_format
    key = (value.shape, value.dtype, value.tobytes(), shape, dtype, device, memory_format)
    t = _constant_cache.get(key)
    if t is None:
        t = torch.as_tensor(value.copy(), dtype=dtype, device=device)
        if shape is not None:
            t, _ = torch.broadcast_tensors(t, torch.empty(shape))
        t = t.contiguous(memory_format=memory_format)
        _constant_cache[key] = t
    return t


nan_to_num = torch.nan_to_num


@contextlib.contextmanager
def suppress_tracer_warnings():
    yield


def assert_shape(tensor, ref_shape):
    """misc.py:84-98: None entries are wildcards."""
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is not None and int(size) != int(ref):
            raise AssertionError(f'Wrong size for dimension {idx}: got {size}, expected {ref}')


def profiled_function(fn):
    """misc.py:102-107: wrap in a torch.autograd.profiler.record_function scope named after the function."""
    def decorator(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    decorator.__name__ = fn.__name__
    return decorator


def params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False, print_=True):
    """misc.py:157-185: copy by name; with require_all every destination tensor must exist in the source with a
    compatible shape (AssertionError otherwise), without it mismatches are skipped silently."""
    assert isinstance(src_module, torch.nn.Module) and isinstance(dst_module, torch.nn.Module)
    src = dict(named_params_and_buffers(src_module))
    with torch.no_grad():
        for name, tensor in named_params_and_buffers(dst_module):
            if name not in src:
                if require_all:
                    print('NotIn src_module', name)
                    raise AssertionError(f'copy_params_and_buffers: {name} missing from the source module')
                continue
            try:
                tensor.copy_(src[name].detach()).requires_grad_(tensor.requires_grad)
            except RuntimeError:
                if require_all:
                    print(name, src[name].shape, tensor.shape)
                    raise AssertionError(f'copy_params_and_buffers: shape mismatch for {name}')
    # the "reload" idiom of the scripts (Cls(*G.init_args, **G.init_kwargs) + copy_params_and_buffers(G, G_new, require_all=True),
    # reenact_avatar_next3d.py:158-160) keeps the checkpoint identity: the pack cache of runtime.prepack stays keyed by it
    if require_all and '_ia_source_hash' in src_module.__dict__:
        dst_module.__dict__['_ia_source_hash'] = src_module.__dict__['_ia_source_hash']


@contextlib.contextmanager
def ddp_sync(module, sync):
    assert isinstance(module, torch.nn.Module)
    if sync or not isinstance(module, torch.nn.parallel.DistributedDataParallel):
        yield
    else:
        with module.no_sync():
            yield


# ---- dnnlib/util.py --------------------------------------------------------------------------------------------------
def get_obj_by_name(name):
    """'pkg.mod.Obj' -> object; the longest importable prefix is the module (dnnlib/util.py:238-292)."""
    parts = name.split('.')
    for i in range(len(parts) - 1, 0, -1):
        try:
            obj = importlib.import_module('.'.join(parts[:i]))
        except ImportError:
            continue
        try:
            for p in parts[i:]:
                obj = getattr(obj, p)
            return obj
        except AttributeError:
            continue
    raise ImportError(name)


def call_func_by_name(*args, func_name=None, **kwargs):
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn)
    return fn(*args, **kwargs)


def construct_class_by_name(*args, class_name=None, **kwargs):
    return call_func_by_name(*args, func_name=class_name, **kwargs)


def is_url(obj, allow_file_urls=False):
    if not isinstance(obj, str) or '://' not in obj:
        return False
    if allow_file_urls and obj.startswith('file://'):
        return True
    return bool(re.match(r'^[a-z][a-z0-9+.-]*://[^/\s]+', obj))


def open_url(url, cache_dir=None, num_attempts=10, verbose=True, return_filename=False, cache=True):
    """Local paths and file:// URLs only (dnnlib/util.py:398-477 also downloads http(s); this build never opens the network)."""
    assert isinstance(url, str)
    if url.startswith('file://'):
        url = url[len('file://'):]
        if re.match(r'^/[a-zA-Z]:', url):
            url = url[1:]
    if '://' in url:
        raise IOError(f'open_url: remote URLs are not supported in this build ({url}); download the file and pass its path')
    return url if return_filename else open(url, 'rb')


# ---- legacy.py -------------------------------------------------------------------------------------------------------
class _Unpickler(pickle.Unpickler):
    """Resolve reference module paths onto this package when the reference tree is not importable."""

    def find_class(self, module, name):
        if module == 'torch_utils.persistence' and name == '_reconstruct_persistent_obj':
            return persistence._reconstruct_persistent_obj
        if module == 'dnnlib.util' and name == 'EasyDict':
            return EasyDict
        return super().find_class(module, name)


def load_network_pkl(f, force_fp16=False):
    """legacy.load_network_pkl (legacy.py:24-60) for PyTorch pickles written by the reference's persistence layer."""
    # content hash of the checkpoint: key of the on-disk cache of packed tensor-core weights (runtime.prepack)
    import hashlib
    import io
    raw = f.read()
    digest = hashlib.sha256(raw).hexdigest()
    data = _Unpickler(io.BytesIO(raw)).load()
    if not isinstance(data, dict):
        raise IOError('load_network_pkl: TensorFlow-era pickles are not supported by this build')
    data.setdefault('training_set_kwargs', None)
    data.setdefault('augment_pipe', None)
    assert isinstance(data['G'], torch.nn.Module)
    if force_fp16:
        raise NotImplementedError('force_fp16: the generator runs fp32 semantics with split-bf16 tensor-core operands on this engine')
    for k in ('G', 'G_ema', 'D'):
        if isinstance(data.get(k), torch.nn.Module):
            data[k].__dict__['_ia_source_hash'] = f'{digest}:{k}'
    return data


# ---- training_avatar_texture/camera_utils.py ----------------------------------------------------------------------------
def create_cam2world_matrix(forward_vector, origin):
    """camera_utils.py:118-137."""
    fwd = forward_vector / forward_vector.norm(dim=-1, keepdim=True)
    up = torch.tensor([0, 1, 0], dtype=torch.float, device=origin.device).expand_as(fwd)
    right = -torch.cross(up, fwd, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    up = torch.cross(fwd, right, dim=-1)
    up = up / up.norm(dim=-1, keepdim=True)
    rot = torch.eye(4, device=origin.device).unsqueeze(0).repeat(fwd.shape[0], 1, 1)
    rot[:, :3, :3] = torch.stack((right, up, fwd), axis=-1)
    trans = torch.eye(4, device=origin.device).unsqueeze(0).repeat(fwd.shape[0], 1, 1)
    trans[:, :3, 3] = origin
    cam2world = (trans @ rot)[:, :, :]
    assert cam2world.shape[1:] == (4, 4)
    return cam2world


def _origins(h, v, radius, device):
    v = torch.clamp(v, 1e-5, math.pi - 1e-5)
    theta = h
    phi = torch.arccos(1 - 2 * (v / math.pi))
    o = torch.zeros((h.shape[0], 3), device=device)
    o[:, 0:1] = radius * torch.sin(phi) * torch.cos(math.pi - theta)
    o[:, 2:3] = radius * torch.sin(phi) * torch.sin(math.pi - theta)
    o[:, 1:2] = radius * torch.cos(phi)
    return o


class GaussianCameraPoseSampler:
    @staticmethod
    def sample(horizontal_mean, vertical_mean, horizontal_stddev=0, vertical_stddev=0, radius=1, batch_size=1, device='cpu'):
        h = torch.randn((batch_size, 1), device=device) * horizontal_stddev + horizontal_mean
        v = torch.randn((batch_size, 1), device=device) * vertical_stddev + vertical_mean
        o = _origins(h, v, radius, device)
        return create_cam2world_matrix(-o, o)


class LookAtPoseSampler:
    @staticmethod
    def sample(horizontal_mean, vertical_mean, lookat_position, horizontal_stddev=0, vertical_stddev=0, radius=1, batch_size=1, device='cpu'):
        h = torch.randn((batch_size, 1), device=device) * horizontal_stddev + horizontal_mean
        v = torch.randn((batch_size, 1), device=device) * vertical_stddev + vertical_mean
        o = _origins(h, v, radius, device)
        return create_cam2world_matrix(lookat_position - o, o)


class UniformCameraPoseSampler:
    @staticmethod
    def sample(horizontal_mean, vertical_mean, horizontal_stddev=0, vertical_stddev=0, radius=1, batch_size=1, device='cpu'):
        h = (torch.rand((batch_size, 1), device=device) * 2 - 1) * horizontal_stddev + horizontal_mean
        v = (torch.rand((batch_size, 1), device=device) * 2 - 1) * vertical_stddev + vertical_mean
        o = _origins(h, v, radius, device)
        return create_cam2world_matrix(-o, o)


def FOV_to_intrinsics(fov_degrees, device='cpu'):
    return synth.fov_to_intrinsics(fov_degrees).to(device)


# ---- reenact_avatar_next3d.py:117-131 ------------------------------------------------------------------------------------
def layout_grid(img, grid_w=None, grid_h=1, float_to_uint8=True, chw_to_hwc=True, to_numpy=True):
    """Output stage of the inference scripts.  The common configuration (uint8 + HWC) is one fused kernel on the device."""
    from . import runtime as rt
    if float_to_uint8 and chw_to_hwc:
        out = rt.layout_grid_u8(img, grid_w=grid_w, grid_h=grid_h)
        return out.cpu().numpy() if to_numpy else out
    raise NotImplementedError('layout_grid: only float_to_uint8=True, chw_to_hwc=True (the configuration the scripts use) is implemented')
