"""
Path statistics of retained draws (oracle; test infrastructure only -- see oracle/__init__.py).

Restates the one-liners the upstream scripts apply to a (n, d) sample array: mode populations and
their KL divergence to the mixture weights (scripts/vMF_diagnostics.py:331-342), the Bingham
hopping frequency (scripts/bingham.py:23-25) and the geodesic distance of successive draws
(geosss/sphere.py:64-68 via scripts/curve_vMF.py:379).  PINNED: tests/golden/pathstats.npz holds the
values the reference expressions give on a reference trajectory (oracle/make_golden.py).
"""

import numpy as np


def mode_populations(x, modes):
    """Draws per mode, label = argmax_k x . mode_k (vMF_diagnostics.py:333-337)."""
    labels = np.argmax(np.asarray(x) @ np.asarray(modes).T, axis=1)
    return np.bincount(labels, minlength=len(modes)).astype(float)


def mode_kl(x, modes, weights):
    """KL(populations || weights), empty modes kept at 1e-100 (vMF_diagnostics.py:338-342)."""
    c = mode_populations(x, modes)
    p = np.full(len(modes), 1e-100)
    p[c > 0] = c[c > 0]
    p[c > 0] /= p.sum()
    return float(p @ np.log(p / np.asarray(weights)))


def hops(x, mode):
    """Successive draws whose sign of x . mode differs (bingham.py:25, before the mean)."""
    return int(np.sum(np.diff(np.sign(np.asarray(x) @ np.asarray(mode))) != 0.0))


def geodesic_steps(x):
    """arccos(clip(x_t . x_{t+1}, -1, 1)) for successive draws (sphere.py:64-68)."""
    x = np.asarray(x)
    return np.arccos(np.clip(np.sum(x[:-1] * x[1:], axis=-1), -1, 1))
