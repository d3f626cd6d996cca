"""
Call counters and the launcher facade (geosss/utils.py:137-232).

The reference decorates ``log_prob`` so that scripts can read ``pdf.log_prob.num_calls`` and
call ``pdf.log_prob.reset_counters()`` (scripts/mixture_vMF.py:107-120).  ``counted`` gives a
method the same two attributes, shared by all instances of the class like upstream's
class-level wrapper.  The device samplers evaluate densities inside kernels, so they add the
number of evaluations the reference would have made (2 + rejections per transition).
"""

import functools


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
