"""
Exact i.i.d. samplers the reference uses as ground truth, generated on the GPU (geosss/rand.py).

    sample_vMF(pdf, size)                 Wood (1994) rejection sampler             rand.py:31-66
    sample_bingham(A, n_samples, ...)     Kent-Ganeiber-Mardia ACG rejection        rand.py:132-193
    sample_bingham_2d / sample_bingham_3d same law through the ACG sampler          rand.py:69-129

Same call signatures and return shapes as upstream, plus ``seed`` (upstream draws from numpy's global
state; here draw ``i`` is a function of ``(seed, i)``), ``dtype`` and ``device``.  ``seed=None`` takes
a fresh seed from numpy's global generator, so ``np.random.seed`` still controls reproducibility.
"""

import ctypes as C

import numpy as np
import torch
from scipy.optimize import brentq

from . import _device, _lib
from .distributions import Bingham, VonMisesFisher


def _seed(seed):
    return int(np.random.randint(0, 2**63 - 1, dtype=np.int64)) if seed is None else _device.seed_to_u64(seed)


def _finish(out, as_torch):
    return out if as_torch else out.cpu().numpy()


def sample_vMF(pdf, size=1, *, seed=None, dtype="float64", device=None, as_torch=False):
    """``size`` draws from vMF(pdf.mu): ``(size, d)``, or ``(d,)`` for ``size == 1`` as upstream."""
    assert isinstance(pdf, VonMisesFisher)
    size = int(size)
    code, tdt, _ = _device.dtype_code(dtype)
    dev = _device.resolve_device(device)
    mu = np.ascontiguousarray(pdf.mu, dtype=float)
    out = torch.empty((size, mu.size), dtype=tdt, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().geosss_exact_vmf(mu.size, _lib.dptr(mu), _seed(seed), 0, code, out.data_ptr(), size, None,
                                               _device.stream_ptr(dev)))
    res = _finish(out, as_torch)
    return res[0] if size == 1 else res


def bfind(vals):
    """Root b of ``sum_j 1 / (b + 2 v_j) = 1`` for the shifted eigenvalues ``vals`` (rand.py:132-142)."""
    vals = np.asarray(vals, dtype=float)
    if np.allclose(vals, 0.0):
        return len(vals)
    return brentq(lambda b: 1 - np.sum(1 / (b + 2 * vals)), 1, len(vals))


def sample_bingham(A, n_samples, n_iter=1000, return_efficiency=False, *, seed=None, dtype="float64", device=None, as_torch=False):
    """``n_samples`` draws from Bingham(A) ~ exp(x^T A x); ``A`` is the matrix or the vector of its diagonal.
    Every draw runs its own rejection loop of at most ``n_iter`` proposals; the efficiency returned on request
    is accepted / proposed over the whole batch."""
    A = np.asarray(A, dtype=float)
    if A.ndim == 1:
        v, U = -A, None
    elif A.ndim == 2:
        v, U = np.linalg.eigh(-A)
        U = np.ascontiguousarray(U)
    else:
        raise ValueError("A must be a matrix or the vector of its diagonal")
    v = np.ascontiguousarray(v - np.min(v))
    d = len(v)
    b = float(bfind(v))
    log_m = -(d - b) / 2 + (d / 2) * np.log(d / b)
    n_samples = int(n_samples)
    code, tdt, _ = _device.dtype_code(dtype)
    dev = _device.resolve_device(device)
    out = torch.empty((n_samples, d), dtype=tdt, device=dev)
    tried = C.c_uint64(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().geosss_exact_bingham(d, _lib.dptr(v), None if U is None else _lib.dptr(U), b, log_m, _seed(seed), 0,
                                                   code, out.data_ptr(), n_samples, int(n_iter), C.byref(tried),
                                                   _device.stream_ptr(dev)))
    res = _finish(out, as_torch)
    return (res, n_samples / max(tried.value, 1)) if return_efficiency else res


def sample_bingham_2d(pdf, n_samples=1, **kwargs):
    """2-D Bingham draws (upstream maps a von Mises variate, rand.py:69-90; the ACG sampler gives the same law)."""
    if not (isinstance(pdf, Bingham) and pdf.d == 2):
        raise ValueError("expected 2D Bingham distribution")
    return sample_bingham(pdf.A, n_samples, **kwargs)


def sample_bingham_3d(pdf, n_samples=1, **kwargs):
    """3-D Bingham draws (upstream runs a Gibbs sampler, rand.py:93-129; these are exact and independent)."""
    if not isinstance(pdf, Bingham) or pdf.d != 3:
        raise ValueError("3D Bingham expected")
    return sample_bingham(pdf.A, n_samples, **kwargs)
