"""invertavatar_b200: B200-native (sm_100a) implementation of the InvertAvatar generator-forward hot path.

Host side: Python modules mirroring the reference's classes (``TriPlaneGenerator`` etc.); device side: hand-written CUDA
behind the C-ABI in ``include/invertavatar_b200.h`` (``libinvertavatar_b200.so``).  There is no CPU fallback."""
__version__ = '0.1.0'


def prepack(module, source_hash=None, cache_dir=None):
    """Pack every convolution weight of ``module`` (already on the GPU) into the tensor-core layout now; with a checkpoint hash
    (recorded by legacy.load_network_pkl) the packs are cached on disk and read back on later loads (see runtime.prepack)."""
    from . import runtime
    return runtime.prepack(module, source_hash=source_hash, cache_dir=cache_dir)
