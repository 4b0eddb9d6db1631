"""CPU: the C-ABI shared library builds, loads and exports every symbol include/invertavatar_b200.h declares
(no compute calls -- there is no GPU here), and the ctypes structures match the header's field lists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'invertavatar_b200.h')


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ia_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from invertavatar_b200 import build, _C
    build.build()
    lib = ctypes.CDLL(_C.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
        assert n in _C.SIGNATURES, f'{n} has no ctypes signature in _C.py'
    for n in _C.SIGNATURES:
        assert n in names, f'{n} bound in _C.py but not declared in the header'


def test_abi_version_and_error_channel():
    from invertavatar_b200 import _C
    lib = _C.lib()
    assert lib.ia_abi_version() == _C.ABI_VERSION
    # argument validation happens before any CUDA call, so it is testable without a device
    rc = lib.ia_bias_act(None, None, None, 4, 1, 1, 1, 0.0, 1.0, -1.0, None)
    assert rc != 0 and b'null' in lib.ia_last_error()
    with pytest.raises(RuntimeError):
        _C.check(rc, 'ia_bias_act')


def _struct_fields(name):
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    structs = {n: b for b, n in re.findall(r'typedef struct \{([^}]*)\}\s*(\w+)\s*;', src)}
    body = structs[name]
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        # "const float* a; int32_t B, H" style: type then comma-separated declarators
        parts = decl.split(',')
        first = parts[0].split()
        fields.append(first[-1].lstrip('*'))
        for p in parts[1:]:
            fields.append(p.strip().lstrip('*'))
    return [re.sub(r'\[.*\]', '', f) for f in fields]


@pytest.mark.parametrize('cname,pyname', [('ia_upfirdn2d_params', 'Upfirdn2dParams'), ('ia_style_layer', 'StyleLayer'),
                                          ('ia_modsplit_params', 'ModsplitParams'), ('ia_emit', 'Emit'), ('ia_conv_params', 'ConvParams'),
                                          ('ia_fir_params', 'FirParams'), ('ia_torgb_params', 'TorgbParams'),
                                          ('ia_resize_params', 'ResizeParams'), ('ia_lerp_params', 'LerpParams'),
                                          ('ia_render_params', 'RenderParams'), ('ia_view', 'View'), ('ia_raster_level_params', 'RasterLevelParams'),
                                          ('ia_enc_prep_params', 'EncPrepParams'), ('ia_enc_affine_params', 'EncAffineParams')])
def test_ctypes_structs_follow_header(cname, pyname):
    from invertavatar_b200 import _C
    want = _struct_fields(cname)
    got = [f[0] for f in getattr(_C, pyname)._fields_]
    got = ['in' if g == 'inp' else g for g in got]
    assert got == want, (cname, got, want)


def test_product_has_no_cpu_fallback():
    import torch
    from invertavatar_b200 import runtime as rt
    with pytest.raises(RuntimeError):
        rt.bias_act(torch.zeros(2, 3), None)
    with pytest.raises(RuntimeError):
        rt.fully_connected(torch.zeros(2, 3), torch.zeros(4, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'invertavatar_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{f} imports the oracle'
