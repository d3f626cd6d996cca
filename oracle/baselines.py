"""
Baseline samplers of the paper's four-way comparison, one chain at a time (oracle; test
infrastructure only -- see oracle/__init__.py).

Restates geosss/mcmc.py:78-112 (burn-in step-size adaptation), :115-173 (random-walk
Metropolis-Hastings), :176-179 (independence sampler), :182-234 (RWMH / independence mixture)
and :237-332 (spherical HMC).  ``rng`` needs ``gamma(shape)``, ``standard_normal(n)`` and
``random()`` -- a numpy Generator, ``tape.TapeRNG`` or ``philox.PhiloxSeqRNG``.

Kept on purpose: the proposal radius is ``sqrt(2 Gamma(d/2))`` (a chi_d variate); acceptance is
``log(U) < delta`` (strict); the step size is multiplied by 1.02 / 0.98 after each of the first
``burnin`` transitions; in the mixture only random-walk proposals adapt; HMC keeps (x, v) as one
state, refreshes v by projecting a normal draw, and radially projects only the final position.
PINNED: tests/golden/baselines.npz (reference trajectories with recorded draws).
"""

import numpy as np

from . import sphere
from .samplers import burnin_count

RWMH, INDEPENDENCE, MIXTURE, HMC = 2, 3, 4, 5


def tangent(v, x):
    """v - x (x . v): mcmc.py:237-241 (x is NOT normalised first)."""
    return v - x * (x @ v)


class Baseline:
    def __init__(self, kind, target, initial_state, rng, stepsize=None, mixing_probability=0.5, n_steps=10):
        self.kind, self.target, self.rng = kind, target, rng
        self.x = np.array(initial_state, dtype=float)
        self.v = np.zeros_like(self.x)  # HMC momentum (mcmc.py:268)
        self.stepsize = float(stepsize if stepsize is not None else (1e-3 if kind == HMC else 1e-1))
        self.alpha, self.n_steps = float(mixing_probability), int(n_steps)
        self.n_accept = self.rwmh_counter = self.indep_counter = 0
        self.reset(0)

    def reset(self, burnin):
        self._counter, self._burnin = 0, int(burnin)

    def _adapt(self, accepted):
        if self._counter < self._burnin:
            self.stepsize *= 1.02 if accepted else 0.98
            self._counter += 1

    # --- proposals -------------------------------------------------------------------------
    def _walk(self):
        d = len(self.x)
        r = np.sqrt(2 * self.rng.gamma(d / 2))
        return sphere.unit(r * self.x + self.stepsize * self.rng.standard_normal(d))

    def _uniform(self):
        return sphere.unit(self.rng.standard_normal(len(self.x)))

    def _leapfrog(self):
        grad, eps = self.target.gradient, self.stepsize
        x = self.x.copy()
        v = tangent(self.rng.standard_normal(len(x)), x)
        self.v = v.copy()
        v = v + 0.5 * eps * tangent(grad(x), x)
        for t in range(self.n_steps):
            nv = np.linalg.norm(v)
            y = x.copy()
            x = y * np.cos(eps * nv) + (v / nv) * np.sin(eps * nv)
            v = v * np.cos(eps * nv) - (y * nv) * np.sin(eps * nv)
            v = v + (eps if t < self.n_steps - 1 else 0.5 * eps) * tangent(grad(x), x)
        return sphere.unit(x), v

    def __next__(self):
        lp = self.target.log_prob
        if hasattr(self.rng, "begin_step"):
            self.rng.begin_step()
        if self.kind == HMC:
            x, v = self._leapfrog()
            delta = (0.5 * self.v @ self.v - lp(self.x)) - (0.5 * v @ v - lp(x))
            accepted = np.log(self.rng.random()) < delta
            self.n_accept += int(accepted)
            self._adapt(accepted)
            if accepted:
                self.x, self.v = x, v
            return self.x
        walk = True
        if self.kind == MIXTURE:
            walk = self.rng.random() < self.alpha
            self.rwmh_counter += int(walk)
            self.indep_counter += int(not walk)
        elif self.kind == INDEPENDENCE:
            walk = False
        y = self._walk() if walk else self._uniform()
        accepted = np.log(self.rng.random()) < lp(y) - lp(self.x)
        self.n_accept += int(accepted)
        if walk or self.kind == INDEPENDENCE:
            self._adapt(accepted)
        if accepted:
            self.x = y
        return self.x

    def sample(self, n_samples, burnin=0, return_all_samples=False):
        """mcmc.py:164-173 on top of :55-77: the first retained state is the initial one."""
        b = burnin_count(n_samples, burnin)
        self.reset(b)
        out = [self.x.copy()]
        while len(out) < n_samples + b:
            out.append(np.copy(next(self)))
        return np.array(out) if return_all_samples else np.array(out[b:])
