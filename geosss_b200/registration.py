"""
Rigid 3D-3D registration likelihoods on S^3 (geosss/registration.py:19-259).

``log_prob`` takes a unit quaternion (x, y, z, w) -- or rows of them, or a 3x3 rotation
matrix -- and evaluates the coherent-point-drift log-likelihood on the GPU: the source cloud
is rotated, every target point sees a Gaussian mixture over its k nearest rotated source
points plus (CPD) a uniform outlier term.  The reference rebuilds a KDTree per call; the
kernels evaluate all pairs and select the k nearest per row exactly (csrc/targets.cuh).
Translations and the 3D->2D projection model are outside this engine's scope.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .distributions import Distribution
from .pointcloud import PointCloud, matrix2quat
from .utils import counted


class Registration(Distribution):
    beta = 1.0
    d = 4

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
        """``dense=True`` sums every source point instead of the k nearest (an extension: it is what the
        float32 throughput kernels evaluate; at the shipped sigma = 1 it equals k = 20 to 5e-12, tested)."""
        if target.dim != 3 or source.dim != 3:
            raise NotImplementedError("the device engine handles 3D-3D registration")
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
        _lib.check(lib.geosss_target_create_cpd(len(tgt), len(src), _lib.dptr(tgt), _lib.dptr(wt), _lib.dptr(src),
                                                _lib.dptr(ws), self.sigma, self.omega, self.beta,
                                                0 if self.dense else self.k, self._outlier, C.byref(h)))
        return h

    @counted
    def log_prob(self, rotation, translation=None):
        if translation is not None:
            raise NotImplementedError("translations are not sampled by the reference's experiments")
        return self._evaluate(self._quaternions(rotation))

    @counted
    def gradient(self, rotation, translation=None):
        """d log_prob / d q for scalar-last quaternions (registration.py:55-59, pointcloud.py:135-203)."""
        if translation is not None:
            raise NotImplementedError("translations are not sampled by the reference's experiments")
        return self._evaluate(self._quaternions(rotation), gradient=True)


class CoherentPointDrift(GaussianMixtureModel):
    """GMM plus a uniform outlier component of weight omega (registration.py:166-259)."""

    _outlier = 1

    def __init__(self, target, source, sigma=1.0, k=20, *, beta=1.0, omega=0.0, dense=False):
        super().__init__(target, source, sigma, k, beta=beta, dense=dense)
        self.omega = float(omega)
        self.log_eps = 1e-308

    @counted
    def log_prob(self, rotation, translation=None):
        if translation is not None:
            raise NotImplementedError("translations are not sampled by the reference's experiments")
        return self._evaluate(self._quaternions(rotation))

    @counted
    def gradient(self, rotation, translation=None):
        """d log_prob / d q for scalar-last quaternions (registration.py:55-59, pointcloud.py:135-203)."""
        if translation is not None:
            raise NotImplementedError("translations are not sampled by the reference's experiments")
        return self._evaluate(self._quaternions(rotation), gradient=True)
