"""Pickle-with-source support, API-compatible with reference ``torch_utils/persistence.py`` (written from
scratch against its behaviour).

A class decorated with ``@persistent_class`` records its constructor arguments (``init_args`` /
``init_kwargs``) and pickles itself as ``(_reconstruct_persistent_obj, (meta,))`` where ``meta`` carries the
source text of the defining module, so the object can be rebuilt where the module is absent.  Reference
pickles use the same record layout (``type='class'``, ``version``, ``module_src``, ``class_name``, ``state``;
persistence.py:120-128,181-206), so ``legacy.load_network_pkl`` can load them through this module, and the
scripts' "reload" idiom ``Cls(*obj.init_args, **obj.init_kwargs)`` + ``copy_params_and_buffers`` works."""
import copy
import importlib
import inspect
import io
import pickle
import re
import sys
import types
import uuid

_VERSION = 6
_decorated = set()
_import_hooks = []
_src_by_module = {}
_module_by_src = {}


class EasyDict(dict):
    """Attribute-style dict (reference dnnlib/util.py EasyDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


def is_persistent(obj):
    try:
        if obj in _decorated:
            return True
    except TypeError:
        pass
    return type(obj) in _decorated


def import_hook(hook):
    assert callable(hook)
    _import_hooks.append(hook)


def _module_source(module):
    src = _src_by_module.get(module)
    if src is None:
        src = inspect.getsource(module)
        _src_by_module[module] = src
        _module_by_src[src] = module
    return src


def _module_from_source(src, package=None):
    module = _module_by_src.get(src)
    if module is None:
        name = '_imported_module_' + uuid.uuid4().hex
        module = types.ModuleType(name)
        if package:
            # the engine's own modules use relative imports (``from . import runtime``): give the rebuilt module the package
            # they resolve against, otherwise exec() fails with "attempted relative import with no known parent package"
            importlib.import_module(package)
            module.__package__ = package
        sys.modules[name] = module
        _src_by_module[module] = src
        _module_by_src[src] = module
        exec(src, module.__dict__)  # the pickled module text (reference behaviour, persistence.py:216-229)
    return module


def _class_by_name(module_name, class_name, src):
    """The class as importable in THIS process (a fresh process that never imported the defining module included).  Used when
    the recorded module is one of this package's and its source is the one that was pickled -- otherwise the pickled source
    text is authoritative (reference behaviour)."""
    if not module_name or not module_name.split('.')[0] == __name__.split('.')[0]:
        return None
    try:
        module = importlib.import_module(module_name)
    except Exception:
        return None
    cls = getattr(module, class_name, None)
    if not isinstance(cls, type):
        return None
    try:
        if src and _module_source(module) != src:
            return None            # the module changed since the pickle was written: rebuild from the pickled text
    except (OSError, TypeError):
        pass
    return cls


def _assert_pickleable(obj):
    def strip(o):
        if isinstance(o, (list, tuple, set)):
            return [strip(x) for x in o]
        if isinstance(o, dict):
            return [[strip(k), strip(v)] for k, v in o.items()]
        if isinstance(o, (str, int, float, bool, bytes, bytearray)) or o is None:
            return None
        if f'{type(o).__module__}.{type(o).__name__}' in ('numpy.ndarray', 'torch.Tensor', 'torch.nn.parameter.Parameter'):
            return None
        if is_persistent(o):
            return None
        return o
    with io.BytesIO() as f:
        pickle.dump(strip(obj), f)


def persistent_class(orig_class):
    assert isinstance(orig_class, type)
    if is_persistent(orig_class):
        return orig_class
    assert orig_class.__module__ in sys.modules
    orig_module = sys.modules[orig_class.__module__]
    try:
        orig_src = _module_source(orig_module)
    except (OSError, TypeError):
        orig_src = ''

    class Decorator(orig_class):
        _orig_module_src = orig_src
        _orig_class_name = orig_class.__name__
        _orig_module_name = orig_class.__module__

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self._init_args = copy.deepcopy(args)
            self._init_kwargs = copy.deepcopy(kwargs)

        @property
        def init_args(self):
            return copy.deepcopy(self._init_args)

        @property
        def init_kwargs(self):
            return EasyDict(copy.deepcopy(self._init_kwargs))

        def __reduce__(self):
            fields = list(super().__reduce__())
            fields += [None] * max(3 - len(fields), 0)
            if fields[0] is not _reconstruct_persistent_obj:
                # module_name is an addition to the reference's record (ignored by the reference's loader): lets a fresh
                # process find the class by import before falling back to the source text
                meta = dict(type='class', version=_VERSION, module_src=self._orig_module_src,
                            class_name=self._orig_class_name, state=fields[2], module_name=self._orig_module_name)
                fields[0] = _reconstruct_persistent_obj
                fields[1] = (meta,)
                fields[2] = None
            return tuple(fields)

    Decorator.__name__ = orig_class.__name__
    Decorator.__qualname__ = orig_class.__qualname__
    Decorator.__module__ = orig_class.__module__
    _decorated.add(Decorator)
    return Decorator


def _reconstruct_persistent_obj(meta):
    meta = EasyDict(meta)
    meta.state = EasyDict(meta.state)
    for hook in _import_hooks:
        meta = hook(meta)
        assert meta is not None
    assert meta.version == _VERSION
    assert meta.type == 'class'
    module_name = meta.get('module_name')
    orig_class = _class_by_name(module_name, meta.class_name, meta.module_src)
    if orig_class is None:
        pkg = module_name.rsplit('.', 1)[0] if (module_name and '.' in module_name and module_name.split('.')[0] == __name__.split('.')[0]) else None
        if pkg is None and not module_name and re.search(r'^from \.+ ?import |^from \.\w', meta.module_src or '', re.M):
            pkg = __name__.rsplit('.', 1)[0]      # a pickle of this engine written before module_name was recorded
        module = _module_from_source(meta.module_src, package=pkg)
        orig_class = module.__dict__[meta.class_name]
    decorator_class = persistent_class(orig_class)
    obj = decorator_class.__new__(decorator_class)
    setstate = getattr(obj, '__setstate__', None)
    if callable(setstate):
        setstate(meta.state)
    else:
        obj.__dict__.update(meta.state)
    return obj
