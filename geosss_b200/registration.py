"""
Rigid 3D-3D and 3D-2D registration likelihoods on S^3 (geosss/registration.py:19-293).

``log_prob`` takes a unit quaternion (x, y, z, w) -- or rows of them, or a 3x3 rotation
matrix -- and evaluates the coherent-point-drift log-likelihood on the GPU: the source cloud
is rotated, every target point sees a Gaussian mixture over its k nearest rotated source
points plus (CPD) a uniform outlier term.  The reference rebuilds a KDTree per call.  The float64 /
literal kernels evaluate all pairs and select the k nearest per row exactly (csrc/targets.cuh); the
float32 throughput kernels of CPD targets (csrc/kernels_cpd_cell.cuh) rotate the target rows into the
source frame and read candidate lists from a grid built once, keeping the k-nearest semantics.
A 2-D target with a ``RotationProjection`` source scores the rotated source projected onto its
first two coordinates (pointcloud.py:273-293).  ``translation`` is a fixed shift added after the
rotation, per call as upstream or persistently through the ``translation`` attribute (which the
samplers then see).
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .distributions import Distribution
from .pointcloud import PointCloud, RotationProjection, matrix2quat
from .utils import counted


class Registration(Distribution):
    beta = 1.0
    d = 4
    translation = None  # fixed shift applied after the rotation; None = no shift

    def _handle(self, dev):
        h = super()._handle(dev)
        t = None if self.translation is None else np.ascontiguousarray(self.translation, dtype=float)
        if t is not None and t.shape != (self.target.dim,):
            raise ValueError(f"translation must have shape ({self.target.dim},)")
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().geosss_target_cpd_set_translation(h, None if t is None else _lib.dptr(t)))
        return h

    def _shifted(self, translation, rotation, gradient):
        """Evaluate with the per-call ``translation`` of upstream's signature (registration.py:47-59)."""
        if translation is None:
            return self._evaluate(self._quaternions(rotation), gradient=gradient)
        keep = self.translation
        self.translation = translation
        try:
            return self._evaluate(self._quaternions(rotation), gradient=gradient)
        finally:
            self.translation = keep

    def _quaternions(self, rotation):
        if isinstance(rotation, torch.Tensor):
            if tuple(rotation.shape) == (3, 3):
                return torch.as_tensor(matrix2quat(rotation.detach().cpu().numpy()), dtype=rotation.dtype)
            return rotation
        rotation = np.asarray(rotation)
        if rotation.shape == (3, 3):
            return matrix2quat(rotation)
        return rotation


class GaussianMixtureModel(Registration):
    """Gaussian mixture over the k nearest rotated source points (registration.py:62-119)."""

    _outlier = 0

    def __init__(self, target: PointCloud, source: PointCloud, sigma: float = 1.0, k: int = 20, *, beta: float = 1.0,
                 dense: bool = False):
        """``dense=True`` sums every source point instead of the k nearest (an extension; at the shipped
        sigma = 1 it equals k = 20 to 5e-12, tested)."""
        if source.dim != 3 or target.dim not in (2, 3):
            raise NotImplementedError("the device engine handles 3D-3D and 3D-2D registration")
        if (target.dim == 2) != isinstance(source, RotationProjection):
            raise ValueError("a 2-D target needs a RotationProjection source, a 3-D target a plain PointCloud")
        self.target, self.source = target, source
        self.sigma, self.k, self.beta = float(sigma), int(k), float(beta)
        self.dense = bool(dense)
        self.omega = 0.0

    def _create_handle(self, lib):
        tgt = np.ascontiguousarray(self.target.positions, dtype=float)
        src = np.ascontiguousarray(self.source.positions, dtype=float)
        wt = np.ascontiguousarray(self.target.weights, dtype=float)
        ws = np.ascontiguousarray(self.source.weights, dtype=float)
        h = C.c_void_p()
        create = lib.geosss_target_create_cpd if self.target.dim == 3 else lib.geosss_target_create_cpd_projected
        _lib.check(create(len(tgt), len(src), _lib.dptr(tgt), _lib.dptr(wt), _lib.dptr(src), _lib.dptr(ws), self.sigma,
                          self.omega, self.beta, 0 if self.dense else self.k, self._outlier, C.byref(h)))
        return h

    @counted
    def log_prob(self, rotation, translation=None):
        return self._shifted(translation, rotation, False)

    @counted
    def gradient(self, rotation, translation=None):
        """d log_prob / d q for scalar-last quaternions (registration.py:55-59, pointcloud.py:135-203)."""
        return self._shifted(translation, rotation, True)

    def cutoff_grid_info(self, device=None):
        """Cell lists the float32 throughput kernels use in place of upstream's per-call KD-tree
        (registration.py:101-109): dict(cutoff, mean_candidates, n_cells, n_entries), or None when the
        target has no such lists (no outlier floor, 2-D target, lists too large to pay)."""
        from . import _device
        h = self._handle(_device.resolve_device(device))
        rc, mc, nc, ne = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        has = _lib.lib().geosss_target_cpd_grid_info(h, C.byref(rc), C.byref(mc), C.byref(nc), C.byref(ne))
        if has < 0:
            _lib.check(has)
        if not has:
            return None
        return {"cutoff": rc.value, "mean_candidates": mc.value, "n_cells": nc.value, "n_entries": ne.value}


class CoherentPointDrift(GaussianMixtureModel):
    """GMM plus a uniform outlier component of weight omega (registration.py:166-259)."""

    _outlier = 1

    def __init__(self, target, source, sigma=1.0, k=20, *, beta=1.0, omega=0.0, dense=False):
        super().__init__(target, source, sigma, k, beta=beta, dense=dense)
        self.omega = float(omega)
        self.log_eps = 1e-308

    @counted
    def log_prob(self, rotation, translation=None):
        return self._shifted(translation, rotation, False)

    @counted
    def gradient(self, rotation, translation=None):
        """d log_prob / d q for scalar-last quaternions (registration.py:55-59, pointcloud.py:135-203)."""
        return self._shifted(translation, rotation, True)
