"""Ahead-of-time build of libinvertavatar_b200.so (sm_100a only; replaces the reference's JIT
``torch_utils/custom_ops.get_plugin``, custom_ops.py:61-149).  Run ``python -m invertavatar_b200.build``.

The library is rebuilt only when the content hash of its sources changes (mtimes do not survive the copy to a
GPU box), so a prebuilt in-tree .so is used as is."""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'libinvertavatar_b200.so')
STAMP = LIB + '.srchash'
SOURCES = ['ia_ops.cu', 'ia_modconv.cu', 'ia_fir_tma.cu', 'ia_conv_tc.cu', 'ia_raster.cu', 'ia_render.cu', 'ia_encoder.cu', 'ia_mesh.cu', 'ia_vit.cu']
HEADERS = [os.path.join(CSRC, 'ia_common.cuh'), os.path.normpath(os.path.join(HERE, '..', 'include', 'invertavatar_b200.h'))]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; libinvertavatar_b200.so cannot be built')
    return exe


def _hash(paths):
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return ''


def source_hash():
    return _hash([os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def is_current():
    return os.path.exists(LIB) and _read(STAMP) == source_hash()


def build(force=False, verbose=False):
    if not force and is_current():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        want = _hash([s] + HEADERS)
        if force or not os.path.exists(o) or _read(o + '.srchash') != want:
            jobs.append((src, [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o], o + '.srchash', want))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        if verbose:
            print(' '.join(cmd))

    def compile_one(job):
        _, cmd, stamp, want = job
        run(cmd)
        with open(stamp, 'w') as f:
            f.write(want)

    if jobs:
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace('.cu', '.o')) for s in SOURCES]
    run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs)
    with open(STAMP, 'w') as f:
        f.write(source_hash())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
