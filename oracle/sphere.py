"""
Sphere geometry used by the slice samplers (oracle; test infrastructure only).

Restates geosss/sphere.py:10-33 (projections), :39-50 (sample_sphere) and
:64-68 (distance).  The ``+1e-100`` in the norm and the order
"normalise pole -> subtract component -> normalise" are kept because the
float64 tape-parity tests depend on them.
"""

import numpy as np

TINY = 1e-100  # geosss/sphere.py:14


def unit(x):
    """x / (|x| + 1e-100), row-wise for rank-2 input (sphere.py:10-18)."""
    x = np.asarray(x)
    nrm = np.linalg.norm(x, axis=-1) + TINY
    return x / nrm if x.ndim == 1 else x / nrm[:, None]


def tangent_part(z, pole):
    """Component of z orthogonal to ``pole`` (sphere.py:21-26)."""
    n = unit(pole)
    return z - (z @ n) * n


def great_subsphere_point(z, pole):
    """Unit vector in the orthogonal complement of ``pole`` (sphere.py:29-33)."""
    return unit(tangent_part(z, pole))


def sample_sphere(d=2, size=None, seed=None):
    """Uniform point(s) on S^d, ambient dimension d+1 (sphere.py:39-50)."""
    gen = np.random.default_rng(seed)
    shape = (d + 1,) if size is None else (size, d + 1)
    return unit(gen.standard_normal(shape))


def distance(x, y):
    """Great-circle distance (sphere.py:64-68)."""
    return np.arccos(np.clip(np.sum(x * y, axis=-1), -1, 1))
