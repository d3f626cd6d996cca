"""
Target densities with the reference's constructors and attributes; ``log_prob`` runs on the
GPU through libgeosss_cuda (no CPU fallback).

    Distribution            geosss/distributions.py:16-25
    VonMisesFisher          :117-160          MixtureModel      :209-227
    Bingham / BinghamFisher :36-114           random_bingham    :230-258
    CurvedVonMisesFisher    :261-278

``log_prob(x)`` accepts a point ``(d,)`` or rows ``(N, d)`` as a numpy array or a torch tensor
(host or CUDA) and returns the same kind.  The arithmetic runs in the input's precision:
float64 input takes the float64 "literal" kernels (reference operation order), anything else
float32.  ``gradient`` (used by the spherical HMC baseline) is batched the same way.
"""

from abc import ABC, abstractmethod
import ctypes as C

import numpy as np
import torch
from scipy.special import i0

from . import _device, _lib, sphere
from .spherical_curve import SlerpCurve, SphericalCurve  # noqa: F401
from .utils import counted


class Distribution(ABC):
    """Interface of a target: ``log_prob(x)`` (+ ``gradient`` upstream)."""

    d = None  # ambient dimension, set by subclasses

    @abstractmethod
    def log_prob(self, x):
        ...

    def gradient(self, x):
        """Gradient of ``log_prob`` in the ambient space, batched on the GPU like ``log_prob``."""
        return self._evaluate(x, gradient=True)

    # ---- device handle -------------------------------------------------------------------
    def _create_handle(self, lib):
        raise NotImplementedError

    def _handle(self, dev):
        cache = self.__dict__.setdefault("_handles", {})
        h = cache.get(dev.index)
        if h is None:
            with torch.cuda.device(dev):
                h = cache[dev.index] = self._create_handle(_lib.lib())
        return h

    def _invalidate(self):
        for idx, h in self.__dict__.pop("_handles", {}).items():
            try:
                with torch.cuda.device(idx):
                    _lib.lib().geosss_target_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self._invalidate()

    def __getstate__(self):  # handles are per-process device objects
        st = dict(self.__dict__)
        st.pop("_handles", None)
        return st

    # ---- batched evaluation --------------------------------------------------------------
    def _check_shape(self, shape):
        if len(shape) not in (1, 2) or shape[-1] != self.d:
            raise ValueError(f"expected a point ({self.d},) or rows (N, {self.d}), got {tuple(shape)}")

    def _evaluate(self, x, device=None, gradient=False):
        is_torch = isinstance(x, torch.Tensor)
        if not is_torch:
            x = np.asarray(x)
        self._check_shape(tuple(x.shape))
        if is_torch and x.is_cuda and device is None:
            dev = x.device
        else:
            dev = _device.resolve_device(device)
        is64 = (x.dtype == torch.float64) if is_torch else (x.dtype != np.float32)
        code, tdt, _ = _device.dtype_code("float64" if is64 else "float32")
        single = x.ndim == 1
        xd = _device.to_device(x, tdt, dev).reshape(-1, self.d)
        out = torch.empty(xd.shape if gradient else xd.shape[0], dtype=tdt, device=dev)
        with torch.cuda.device(dev):
            if xd.shape[0] == 0:
                return out if is_torch else out.cpu().numpy()
            lib, h, args = _lib.lib(), self._handle(dev), (xd.data_ptr(), out.data_ptr(), xd.shape[0], _device.stream_ptr(dev))
            if gradient:
                rc = lib.geosss_gradient(h, code, *args)
            else:
                # float32: bandwidth-bound tiled kernels where one exists for the target, else the literal kernel
                rc = lib.geosss_log_prob(h, code, _lib.PATH_FAST, *args) if code == _lib.F32 else _lib.E_UNSUPPORTED
                if rc == _lib.E_UNSUPPORTED:
                    rc = lib.geosss_log_prob(h, code, _lib.PATH_LITERAL, *args)
            _lib.check(rc)
        if is_torch:
            res = out if x.is_cuda else out.cpu()
            return res[0] if single else res
        res = out.cpu().numpy()
        return res[0] if single else res


class VonMisesFisher(Distribution):
    """vMF(x) ~ exp(mu . x); the reference's constant is -log(2 pi) - log(i0(|mu|)) whatever d is."""

    def __init__(self, mu):
        self.mu = np.array(mu, dtype=float)

    @property
    def d(self):
        return self.mu.size

    @property
    def kappa(self):
        return np.linalg.norm(self.mu)

    @property
    def mode(self):
        return sphere.radial_projection(self.mu)

    @property
    def max_log_prob(self):
        return self.kappa

    def _create_handle(self, lib):
        mu = np.ascontiguousarray(self.mu.reshape(1, -1))
        with np.errstate(over="ignore"):
            li0 = np.log(i0(np.linalg.norm(mu, axis=1)))
        h = C.c_void_p()
        _lib.check(lib.geosss_target_create_vmf_mixture(mu.shape[1], 1, _lib.dptr(mu), None, _lib.dptr(li0), C.byref(h)))
        return h

    @counted
    def log_prob(self, x):
        return self._evaluate(x)

    @counted
    def gradient(self, x):
        return self._evaluate(x, gradient=True)


class MixtureModel(Distribution):
    """Mixture of components; weights are normalised, uniform by default (distributions.py:211-216).

    The device engine evaluates mixtures of VonMisesFisher components (what the reference's
    experiments use); other component types raise NotImplementedError.
    """

    def __init__(self, components, weights=None):
        self.pdfs = components
        if weights is None:
            weights = np.ones(len(components))
        self.weights = np.array(weights, dtype=float)
        self.weights /= self.weights.sum()

    @property
    def d(self):
        return self.pdfs[0].d

    def _create_handle(self, lib):
        if not all(isinstance(p, VonMisesFisher) for p in self.pdfs):
            raise NotImplementedError("MixtureModel on the device supports VonMisesFisher components")
        mu = np.ascontiguousarray(np.stack([p.mu for p in self.pdfs]))
        with np.errstate(over="ignore", divide="ignore"):
            li0 = np.log(i0(np.linalg.norm(mu, axis=1)))
            logw = np.log(self.weights)
        h = C.c_void_p()
        _lib.check(lib.geosss_target_create_vmf_mixture(mu.shape[1], mu.shape[0], _lib.dptr(mu), _lib.dptr(logw),
                                                        _lib.dptr(li0), C.byref(h)))
        return h

    @counted
    def log_prob(self, x):
        # upstream evaluates every component through its own counted log_prob (distributions.py:219)
        VonMisesFisher.log_prob.num_calls += len(self.pdfs)
        return self._evaluate(x)

    @counted
    def gradient(self, x):
        """Responsibility-weighted component gradients (distributions.py:223-227)."""
        VonMisesFisher.log_prob.num_calls += len(self.pdfs)
        VonMisesFisher.gradient.num_calls += len(self.pdfs)
        return self._evaluate(x, gradient=True)


class Bingham(Distribution):
    """Bingham(x) ~ exp(x^T A x), A symmetric."""

    def __init__(self, A):
        A = np.asarray(A, dtype=float)
        assert np.allclose(A, A.T)
        self.A = A
        lam, vec = np.linalg.eigh(A)  # ascending
        self.v, self.U = lam[::-1], vec[:, ::-1]

    @property
    def d(self):
        return len(self.A)

    @property
    def mode(self):
        return self.U[:, 0]

    @property
    def max_log_prob(self):
        return self.v[0]

    def _check_shape(self, shape):
        assert len(shape) in (1, 2) and shape[-1] == self.d  # distributions.py:84

    def _fisher(self):
        return None

    def device_eigenrows(self, device=None):
        """Eigenvector rows ``W`` of the coordinates ``y = W x`` in which the float32 kernels evaluate and sample a DENSE
        ``A`` (d <= 64), or None (diagonal ``A``, larger d).  Chains started from the same seed use the seed's normal
        draws in those coordinates, i.e. the direction ``W.T @ z`` in the caller's."""
        dev = _device.resolve_device(device)
        W = np.empty((self.d, self.d))
        with torch.cuda.device(dev):
            rc = _lib.lib().geosss_target_bingham_eigenrows(self._handle(dev), _lib.dptr(W))
        if rc < 0:
            _lib.check(rc)
        return W if rc == 1 else None

    def _create_handle(self, lib):
        A = np.ascontiguousarray(self.A)
        b = self._fisher()
        h = C.c_void_p()
        _lib.check(lib.geosss_target_create_bingham(self.d, _lib.dptr(A), _lib.dptr(b), C.byref(h)))
        return h

    @counted
    def log_prob(self, x):
        return self._evaluate(x)

    @counted
    def gradient(self, x):
        return self._evaluate(x, gradient=True)


class BinghamFisher(Bingham):
    """exp(x^T A x + x . b)  (distributions.py:106-114)."""

    def __init__(self, A, b):
        super().__init__(A)
        assert len(A) == len(b)
        self.b = np.asarray(b, dtype=float)

    def _fisher(self):
        return np.ascontiguousarray(self.b)

    @counted
    def log_prob(self, x):
        return self._evaluate(x)

    @counted
    def gradient(self, x):
        return self._evaluate(x, gradient=True)


def random_bingham(d=2, vmax=None, vmin=None, eigensystem=False, seed=None):
    """Random Bingham target: Wishart-like precision, spectrum shifted to vmin and scaled to vmax,
    optionally in its own eigenbasis (diagonal A)  (distributions.py:230-258)."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((d, d))
    lam, vec = np.linalg.eigh(G.T @ G)
    if vmin is not None:
        lam += vmin - lam.min()
    if vmax is not None:
        lam *= vmax / lam.max()
    if eigensystem:
        vec = np.eye(d)
    return Bingham((vec * lam) @ vec.T)


class CurvedVonMisesFisher(Distribution):
    """exp(kappa * x . nearest_point_on_curve(x)) for a piecewise-slerp curve."""

    def __init__(self, curve, kappa=100.0):
        self.curve = curve
        self.kappa = kappa

    @property
    def d(self):
        return self.curve.knots.shape[-1]

    def _create_handle(self, lib):
        if not isinstance(self.curve, SlerpCurve):
            raise NotImplementedError("the device engine evaluates SlerpCurve targets")
        knots = np.ascontiguousarray(self.curve.knots, dtype=float)
        h = C.c_void_p()
        _lib.check(lib.geosss_target_create_curved_vmf(knots.shape[1], knots.shape[0], _lib.dptr(knots),
                                                       float(self.kappa), C.byref(h)))
        return h

    @counted
    def log_prob(self, x):
        return self._evaluate(x)

    @counted
    def gradient(self, x):
        return self._evaluate(x, gradient=True)
