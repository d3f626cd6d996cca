"""
Host-side sphere helpers with the reference's names and semantics (geosss/sphere.py:10-68).

They are set-up utilities (initial states, fixtures, diagnostics on returned samples); the
per-step geometry of the samplers runs inside the CUDA kernels (csrc/transition.cuh,
csrc/kernels_fast.cuh).  ``sample_sphere`` keeps numpy's PCG64 so that seeds used with the
reference give the same initial states here.
"""

import numpy as np

_EPS = 1e-100


def radial_projection(x):
    """Scale point(s) onto the unit sphere: x / (|x| + 1e-100)  (sphere.py:10-18)."""
    x = np.asarray(x)
    length = np.linalg.norm(x, axis=-1) + _EPS
    return x / (length if x.ndim == 1 else length[:, None])


def orthogonal_projection(x, y):
    """Remove from x its component along y  (sphere.py:21-26)."""
    pole = radial_projection(y)
    return x - (x @ pole) * pole


def spherical_projection(x, v):
    """Unit vector(s) on the great subsphere with pole v  (sphere.py:29-33)."""
    return radial_projection(orthogonal_projection(x, v))


def sample_sphere(d=2, size=None, seed=None):
    """Uniform draw(s) from S^d; note d is the SPHERE dimension, points live in R^(d+1)
    (sphere.py:39-50)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal(d + 1 if size is None else (size, d + 1))
    return radial_projection(z)


def sample_subsphere(v, seed=None):
    """Uniform draw from the great subsphere orthogonal to v  (sphere.py:53-58)."""
    rng = np.random.default_rng(seed)
    return spherical_projection(rng.standard_normal(len(v)), v)


def distance(x, y):
    """Geodesic distance between unit vectors  (sphere.py:64-68)."""
    return np.arccos(np.clip(np.sum(x * y, axis=-1), -1, 1))


# ---- coordinate changes and plane rotations (geosss/sphere.py:74-130); host helpers of the scripts
def cartesian2polar(x):
    """Angle in [0, 2 pi) of points (n, 2) on the circle."""
    x = np.asarray(x)
    return np.mod(np.arctan2(x[:, 1], x[:, 0]), 2 * np.pi)


def polar2cartesian(theta):
    return np.stack([np.cos(theta), np.sin(theta)], axis=-1)


def spherical2cartesian(phi, theta):
    """Azimuth phi, polar angle theta -> points on S^2."""
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1)


def cartesian2spherical(x):
    x = np.asarray(x)
    return cartesian2polar(x), np.mod(np.arccos(x[:, 2]), np.pi)


def givens(u, v, x):
    """Function of theta rotating ``x`` by theta in the plane spanned by the orthonormal pair u, v."""
    ux, vx = u @ x, v @ x

    def rotate(theta):
        return x + (np.cos(theta) - 1) * (ux * u + vx * v) + np.sin(theta) * (ux * v - vx * u)

    return rotate
