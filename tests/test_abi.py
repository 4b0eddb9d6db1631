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
    rc = lib.ia_bias_act(None, None, None, 4, 1, 1, 1, 0.0, 1.0, -1.0, 0, None)
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


@pytest.mark.parametrize('cname,pyname', [('ia_upfirdn2d_params', 'Upfirdn2dParams'), ('ia_filtered_lrelu_params', 'FilteredLreluParams'), ('ia_style_layer', 'StyleLayer'),
                                          ('ia_modsplit_params', 'ModsplitParams'), ('ia_emit', 'Emit'), ('ia_conv_params', 'ConvParams'),
                                          ('ia_fir_params', 'FirParams'), ('ia_torgb_params', 'TorgbParams'),
                                          ('ia_resize_params', 'ResizeParams'), ('ia_lerp_params', 'LerpParams'),
                                          ('ia_render_params', 'RenderParams'), ('ia_view', 'View'), ('ia_raster_level_params', 'RasterLevelParams'),
                                          ('ia_enc_prep_params', 'EncPrepParams'), ('ia_enc_affine_params', 'EncAffineParams'),
                                          ('ia_stitch_params', 'StitchParams'), ('ia_blendshape_params', 'BlendshapeParams'),
                                          ('ia_ortho_raster_params', 'OrthoRasterParams'),
                                          ('ia_enc_im2col_params', 'EncIm2colParams'), ('ia_attention_params', 'AttentionParams'),
                                          ('ia_attention_tc_params', 'AttentionTcParams')])
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


def integration_stub_namespace():
    """Execute, verbatim, the reference-side binding INTEGRATION.md section 3 tells a maintainer to add (the python block that
    starts with `# torch_utils/ops/bias_act.py`).  Its ctypes.CDLL('libinvertavatar_b200.so') resolves through the dynamic
    loader, so the package directory is preloaded by absolute path first (same soname)."""
    from invertavatar_b200 import _C
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
    stub = [b for b in blocks if b.startswith('# torch_utils/ops/bias_act.py')]
    assert len(stub) == 1, 'INTEGRATION.md must hold exactly one bias_act stub block'
    real_cdll = ctypes.CDLL

    def cdll(name, *a, **k):
        return real_cdll(_C.LIB_PATH if name == 'libinvertavatar_b200.so' else name, *a, **k)
    ns = {}
    ctypes.CDLL = cdll
    try:
        exec(compile(stub[0], 'INTEGRATION.md:stub', 'exec'), ns)
    finally:
        ctypes.CDLL = real_cdll
    return ns


def test_integration_stub_argument_validation():
    """The stub runs against the built library without a GPU as far as the library's argument validation: a bad activation
    id, an unsupported dtype and a bias without its channel count all come back as RuntimeError (TORCH_CHECK parity,
    bias_act.cpp:39-55) before any CUDA call is made."""
    import torch
    from types import SimpleNamespace
    ns = integration_stub_namespace()
    fwd = ns['_bias_act_cuda_forward']
    x = torch.zeros(2, 3, 4)
    with pytest.raises(RuntimeError, match='activation'):
        fwd(x, None, 1, SimpleNamespace(cuda_idx=42), 0.0, 1.0, None)
    with pytest.raises(RuntimeError, match='float16, float32 or float64'):
        fwd(x.to(torch.bfloat16), None, 1, SimpleNamespace(cuda_idx=1), 0.0, 1.0, None)
    lib = ns['_lib']
    rc = lib.ia_bias_act(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.numel(), 0, 0, 1, 0.0, 1.0, -1.0, 0, None)
    assert rc != 0 and b'bias needs' in lib.ia_last_error()


def test_filtered_lrelu_export_validates_and_reports_missing_kernel():
    """ia_filtered_lrelu: invalid arguments -> error code + message; sign tensors requested -> -1, the reference plugin's
    "no specialised kernel" code (filtered_lrelu.cpp:56-60) on which its Python falls back (filtered_lrelu.py:225-231)."""
    from invertavatar_b200 import _C
    lib = _C.lib()
    p = _C.FilteredLreluParams()
    assert lib.ia_filtered_lrelu(ctypes.byref(p), None) > 0 and b'null tensor' in lib.ia_last_error()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    p.x, p.y = addr, addr
    p.N, p.C, p.inH, p.inW, p.outH, p.outW = 1, 1, 4, 4, 4, 4
    p.up, p.down, p.gain, p.slope, p.clamp = 1, 1, 1.0, 0.2, -1.0
    assert lib.ia_filtered_lrelu_workspace(ctypes.byref(p)) == 4 * 4 * 4
    p.write_signs = 1
    assert lib.ia_filtered_lrelu(ctypes.byref(p), None) == -1
    p.write_signs = 0
    p.outH = 5
    assert lib.ia_filtered_lrelu(ctypes.byref(p), None) > 0 and b'caller allocated' in lib.ia_last_error()
    p.outH, p.dtype = 4, 2
    assert lib.ia_filtered_lrelu(ctypes.byref(p), None) > 0 and b'float16 or float32' in lib.ia_last_error()
    assert lib.ia_filtered_lrelu_act(addr, 16, 0, addr, 0, 0, 1.0, 0.2, -1.0, 0, None, None) == -1
