"""reference torch_utils/custom_ops.py:61-149 JIT-compiles the plugins; here the library is built ahead of time."""
from invertavatar_b200 import _C, build

verbosity = 'brief'


def get_plugin(module_name=None, sources=None, headers=None, source_dir=None, **build_kwargs):
    build.build()
    return _C.lib()
