"""Compact fingerprints of large tensors, so that golden vectors of 512^2 images and 256^2 x 512-channel
feature maps fit in a few hundred KB of committed fixtures.

A fingerprint keeps (a) a strided subsample of the flattened tensor (exact values, compared element-wise),
(b) float64 mean and mean-|x| and (c) a projection on a seeded +-1 vector (sensitive to a change in ANY element).
"""
import numpy as np

MAX_SUB = 8192


def _signs(n, seed=12345):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 2, size=n).astype(np.float64) * 2.0 - 1.0


def fingerprint(t, max_sub=MAX_SUB):
    """t: torch tensor or ndarray -> dict of small ndarrays."""
    a = t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)
    flat = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    n = flat.size
    step = max(1, -(-n // max_sub))
    # odd step so the subsample does not alias with power-of-two image rows
    if step > 1 and step % 2 == 0:
        step += 1
    f64 = flat.astype(np.float64)
    return {
        'shape': np.asarray(a.shape, dtype=np.int64),
        'step': np.asarray(step, dtype=np.int64),
        'sub': flat[::step].copy(),
        'mean': np.asarray(f64.mean()),
        'absmean': np.asarray(np.abs(f64).mean()),
        'proj': np.asarray(float(np.dot(f64, _signs(n)) / np.sqrt(n))),
    }


def pack(prefix, fp, out):
    for k, v in fp.items():
        out[f'{prefix}/{k}'] = v


def unpack(prefix, npz):
    return {k: npz[f'{prefix}/{k}'] for k in ('shape', 'step', 'sub', 'mean', 'absmean', 'proj')}


def compare(t, fp, atol, name=''):
    """Returns (max_abs_err_on_subsample, dict of scalar errors); raises AssertionError on shape mismatch."""
    got = fingerprint(t)
    assert tuple(got['shape']) == tuple(fp['shape']), f'{name}: shape {tuple(got["shape"])} != golden {tuple(fp["shape"])}'
    assert int(got['step']) == int(fp['step'])
    d = np.abs(got['sub'].astype(np.float64) - fp['sub'].astype(np.float64))
    err = float(d.max()) if d.size else 0.0
    scal = {k: abs(float(got[k]) - float(fp[k])) for k in ('mean', 'absmean', 'proj')}
    assert err <= atol, f'{name}: subsample max-abs error {err:.3e} > {atol:.1e}'
    # mean / absmean are averages of per-element errors, so the same bound applies; the projection is a random-sign
    # sum normalised by sqrt(n): for independent errors of size <= atol it stays below a few atol
    assert scal['mean'] <= atol and scal['absmean'] <= atol, f'{name}: mean/absmean drift {scal} > {atol:.1e}'
    assert scal['proj'] <= 8 * atol, f'{name}: projection drift {scal["proj"]:.3e} > {8 * atol:.1e}'
    return err, scal
