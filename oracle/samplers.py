"""
Geodesic slice samplers, one chain at a time (oracle; test infrastructure only).

Restates geosss/mcmc.py:13-19 (burn-in rule), :28-45 + :55-77 (driver loop) and
:357-374 / :382-401 (reject / shrink transitions).  ``rng`` is anything with
``standard_normal(n)``, ``random()`` and ``uniform(a, b)`` -- a numpy Generator,
a ``tape.TapeRNG`` or a ``philox.PhiloxChainRNG``.

Kept on purpose (float64 tape parity depends on them): the initial state is
neither copied nor normalised; log p(x) is re-evaluated at every transition;
acceptance is a strict ``>``; the shrink side is chosen by ``theta < 0``; the
first retained sample is the initial state, so ``n + burnin - 1`` transitions run.
"""

import numpy as np

from . import sphere

TWO_PI = 2 * np.pi

REJECT, SHRINK = 0, 1


def burnin_count(n_samples, burnin):
    """float in [0,1] -> fraction of n_samples; otherwise a count (mcmc.py:13-19)."""
    if isinstance(burnin, float):
        assert 0 <= burnin <= 1.0
        return int(burnin * n_samples)
    assert burnin >= 0
    return int(burnin)


def transition(kind, log_prob, x, rng):
    """One GeoSSS transition from x. Returns (new state, rejected proposals)."""
    u = sphere.great_subsphere_point(rng.standard_normal(len(x)), x)
    level = log_prob(x) + np.log(rng.random())
    if kind == SHRINK:
        hi = rng.uniform(0, TWO_PI)
        lo = hi - TWO_PI
    else:
        lo, hi = 0, TWO_PI
    rejected = 0
    while True:
        theta = rng.uniform(lo, hi)
        y = np.cos(theta) * x + np.sin(theta) * u
        if log_prob(y) > level:
            return y, rejected
        if kind == SHRINK:
            if theta < 0:
                lo = theta
            else:
                hi = theta
        rejected += 1


class SliceSampler:
    """Driver with the reference's ``sample`` semantics (mcmc.py:55-77)."""

    def __init__(self, kind, log_prob, initial_state, rng):
        self.kind = kind
        self.log_prob = log_prob
        self.state = initial_state
        self.rng = rng
        self.n_reject = 0
        self.rejects_per_step = []

    def __next__(self):
        self.state, r = transition(self.kind, self.log_prob, self.state, self.rng)
        self.n_reject += r
        self.rejects_per_step.append(r)
        return self.state

    def sample(self, n_samples, burnin=0, return_all_samples=False):
        assert n_samples > 0
        b = burnin_count(n_samples, burnin)
        out = [self.state]
        while len(out) < n_samples + b:
            out.append(np.copy(next(self)))
        return np.array(out) if return_all_samples else np.array(out[b:])
