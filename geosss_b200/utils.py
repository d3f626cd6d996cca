"""
Call counters and the launcher facade (geosss/utils.py:137-232).

The reference decorates ``log_prob`` so that scripts can read ``pdf.log_prob.num_calls`` and
call ``pdf.log_prob.reset_counters()`` (scripts/mixture_vMF.py:107-120).  ``counted`` gives a
method the same two attributes, shared by all instances of the class like upstream's
class-level wrapper.  The device samplers evaluate densities inside kernels, so they add the
number of evaluations the reference would have made (2 + rejections per transition).
"""

import contextlib
import functools
import logging
import time

import numpy as np

# the plotting palette of the scripts (utils.py:21-27)
colors = [(0.85, 0.3, 0.1), (0.15, 0.35, 0.6), (0.95, 0.7, 0.1), (0.0, 0.0, 0.0), (0.8, 0.8, 0.8)]


@contextlib.contextmanager
def take_time(desc, mute=False):
    """Log the CPU time of the enclosed block (utils.py:40-46)."""
    t0 = time.process_time()
    yield
    dt = time.process_time() - t0
    if not mute:
        scale, unit = next(((s, u) for s, u in ((1.0, "s"), (1e-3, "ms"), (1e-6, "us")) if dt > s), (1e-9, "ns"))
        logging.info("%s took %.1f %s", desc, dt / scale, unit)


def acf(x, n_max=None):
    """Autocorrelation by direct lagged products, normalised by lag 0 (utils.py:91-106)."""
    x = np.asarray(x, dtype=float)
    n_max = n_max or len(x) // 2
    x = x - x.mean()
    ac = np.array([np.mean(x[i:] * x[:len(x) - i]) for i in range(n_max)])
    return ac / ac[0]


def acf_fft(x):
    """Circular autocorrelation by the convolution theorem (utils.py:109-112)."""
    x = np.asarray(x, dtype=float)
    X = np.fft.rfft((x - x.mean()) / x.std())
    return np.real(np.fft.irfft(X.conj() * X))[:len(x) // 2] / len(x)


def IAT(x, n=None):
    """Integrated autocorrelation time with Gelman's pair-sum truncation (utils.py:115-129)."""
    ac = acf_fft(x)
    if n:
        ac = ac[:n]
    body = ac[2:-1] if len(ac) % 2 != 0 else ac[2:]
    pairs = body.reshape(-1, 2).sum(1)
    neg = np.nonzero(pairs < 0)[0]
    L = 1 + 2 * neg[0] if len(neg) else len(ac) - 1  # upstream indexes neg[0] first and fails without a negative pair
    return 1.0 + max(2 * np.sum(ac[1:L + 1]), 0.0)


def n_eff(x, n=None):
    """Effective sample size len(x) / IAT (utils.py:132-134).  Multi-chain ESS on the GPU: diagnostics.ess."""
    return len(x) / IAT(x, n)


class _BoundCounted:
    def __init__(self, owner, instance):
        self._owner, self._instance = owner, instance

    def __call__(self, *args, **kwargs):
        self._owner.num_calls += 1
        return self._owner.func(self._instance, *args, **kwargs)

    @property
    def num_calls(self):
        return self._owner.num_calls

    def reset_counters(self):
        self._owner.num_calls = 0

    def add_calls(self, n):
        self._owner.num_calls += int(n)


class counted:
    """Descriptor: method with ``.num_calls`` / ``.reset_counters()`` (utils.py:137-160)."""

    def __init__(self, func):
        self.func = func
        self.num_calls = 0
        functools.update_wrapper(self, func)

    def __get__(self, instance, owner=None):
        return self if instance is None else _BoundCounted(self, instance)

    def reset_counters(self):
        self.num_calls = 0


def counter(method_names):
    """Class decorator with upstream's name and arguments (utils.py:162-185): the named methods get
    ``.num_calls`` / ``.reset_counters()``."""
    if isinstance(method_names, str):
        method_names = [method_names]

    def class_decorator(cls):
        for name in method_names:
            raw = cls.__dict__.get(name) or getattr(cls, name, None)
            if raw is not None and not isinstance(raw, counted):
                setattr(cls, name, counted(raw))
        return cls

    return class_decorator


class SamplerLauncher:
    """Method name -> sampler dispatch (utils.py:188-232); keeps the sampler on ``.rsss`` /
    ``.ssss`` / ``.rwmh`` / ``.hmc``."""

    def __init__(self, pdf, initial, n_samples, burnin=0.2, seed=None, **sampler_kwargs):
        self.pdf, self.initial = pdf, initial
        self.n_samples, self.burnin, self.seed = n_samples, burnin, seed
        self.sampler_kwargs = sampler_kwargs

    def run_sss_reject(self):
        from .mcmc import RejectionSphericalSliceSampler

        self.rsss = RejectionSphericalSliceSampler(self.pdf, self.initial, self.seed, **self.sampler_kwargs)
        return self.rsss.sample(self.n_samples, burnin=self.burnin)

    def run_sss_shrink(self):
        from .mcmc import ShrinkageSphericalSliceSampler

        self.ssss = ShrinkageSphericalSliceSampler(self.pdf, self.initial, self.seed, **self.sampler_kwargs)
        return self.ssss.sample(self.n_samples, burnin=self.burnin)

    def run_rwmh(self):
        from .mcmc import MetropolisHastings

        self.rwmh = MetropolisHastings(self.pdf, self.initial, self.seed, stepsize=1e-1, **self.sampler_kwargs)
        return self.rwmh.sample(self.n_samples, burnin=self.burnin)

    def run_hmc(self):
        from .mcmc import SphericalHMC

        self.hmc = SphericalHMC(self.pdf, self.initial, self.seed, stepsize=1e-1, **self.sampler_kwargs)
        return self.hmc.sample(self.n_samples, burnin=self.burnin)

    def run(self, method):
        if method == "sss-reject":
            return self.run_sss_reject()
        if method == "sss-shrink":
            return self.run_sss_shrink()
        if method == "rwmh":
            return self.run_rwmh()
        if method == "hmc":
            return self.run_hmc()
        raise ValueError(f"method {method} not known")
