"""
Curves on the sphere used by the curved-vMF target (geosss/spherical_curve.py:35-129).

Host-side parameter holders: the nearest-point search that the samplers need runs on the GPU
inside CurvedVonMisesFisher's kernels (csrc/targets.cuh CurvedP).  ``find_nearest`` here is a
float64 numpy convenience for single points (plots, checks), written in closed form over all
segments at once.
"""

import numpy as np

from . import sphere


class SphericalCurve:
    def __init__(self, knots):
        self.knots = np.asarray(knots, dtype=float)

    def find_nearest(self, point, *args, **kwargs):
        raise NotImplementedError


class SlerpCurve(SphericalCurve):
    """Piecewise spherical-linear interpolation through the knots (spherical_curve.py:74-102)."""

    def __init__(self, knots):
        super().__init__(knots)
        seg = sphere.distance(self.knots[:-1], self.knots[1:])
        self.theta = np.concatenate([[0.0], seg])
        self.bins = np.cumsum(self.theta)
        self.bins /= self.bins[-1]
        self.widths = np.diff(self.bins)

    def __call__(self, t):
        t = np.atleast_1d(np.asarray(t, dtype=float))
        s = self.bins
        i = np.clip(np.digitize(t, s, right=True), 1, len(self.knots) - 1)
        ang = self.theta[i]
        frac = (t - s[i - 1]) / (s[i] - s[i - 1])
        wa = np.sin(ang * (1 - frac)) / np.sin(ang)
        wb = np.sin(ang * frac) / np.sin(ang)
        return wa[:, None] * self.knots[i - 1] + wb[:, None] * self.knots[i]

    def find_nearest(self, point):
        """Closest curve point to ``point`` (first segment wins ties), spherical_curve.py:10-32,95-102."""
        a, b = self.knots[:-1], self.knots[1:]
        ang = self.theta[1:]
        ga, gb = a @ point, b @ point
        t = np.clip(np.arctan2(gb - ga * np.cos(ang), ga * np.sin(ang)), 0.0, ang)
        wa, wb = np.sin(ang - t), np.sin(t)
        den = np.sin(ang) + 1e-10
        dots = (wa * ga + wb * gb) / den
        j = int(np.argmin(np.arccos(np.clip(dots, -1, 1))))
        return (wa[j] * a[j] + wb[j] * b[j]) / den[j]


def brownian_curve(n_points=100, dimension=6, step_size=0.05, seed=1234):
    """Knots from a Gaussian random walk re-projected onto the sphere after every step
    (spherical_curve.py:105-129); same PCG64 draws as upstream for a given seed."""
    rng = np.random.default_rng(seed)
    knots = np.zeros((n_points, dimension))
    knots[0] = sphere.radial_projection(rng.standard_normal(dimension))
    for i in range(1, n_points):
        knots[i] = sphere.radial_projection(knots[i - 1] + rng.normal(size=dimension) * step_size)
    return knots
