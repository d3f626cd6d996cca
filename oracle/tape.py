"""
Record / replay shim for the RNG calls the samplers make
(oracle; test infrastructure only).

Assigned to ``sampler.rng`` of the upstream samplers it reproduces their own
PCG64 run bitwise (numpy's ``uniform(a, b)`` is ``a + (b - a) * random()``) while
writing every value drawn, in consumption order, to one flat float64 tape:
``d`` normals, the level uniform, [the shrink bracket uniform,] then one uniform
per proposal.  The CUDA tape mode (``geosss_sampler_run_tape``) consumes the same
flat tape.
"""

import numpy as np


class TapeRNG:
    def __init__(self, source=None, tape=None):
        """``source``: numpy Generator to record from; ``tape``: values to replay."""
        assert (source is None) != (tape is None)
        self.source = source
        self.values = [] if tape is None else None
        self.tape = None if tape is None else np.asarray(tape, dtype=float)
        self.cursor = 0

    # --- the three calls geosss/mcmc.py:362-369,387-395 makes -----------------
    def standard_normal(self, n):
        if self.source is not None:
            z = self.source.standard_normal(n)
            self.values.extend(z.tolist())
            return z
        z = self.tape[self.cursor:self.cursor + n].copy()
        assert len(z) == n, "tape exhausted"
        self.cursor += n
        return z

    def random(self):
        if self.source is not None:
            u = self.source.random()
            self.values.append(u)
            return u
        u = self.tape[self.cursor]
        self.cursor += 1
        return u

    def gamma(self, shape):
        """Radius draw of the random-walk proposal (geosss/mcmc.py:142): the gamma variate itself goes on
        the tape."""
        if self.source is not None:
            g = self.source.gamma(shape)
            self.values.append(g)
            return g
        g = self.tape[self.cursor]
        self.cursor += 1
        return g

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def recorded(self):
        return np.array(self.values, dtype=float)
