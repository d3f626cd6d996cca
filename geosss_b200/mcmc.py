"""
Geodesic slice samplers with the reference's interface, running N chains at once on a B200.

    determine_burnin                       geosss/mcmc.py:13-19
    Sampler.__init__/__next__/sample       geosss/mcmc.py:28-77
    RejectionSphericalSliceSampler         geosss/mcmc.py:335-374
    ShrinkageSphericalSliceSampler         geosss/mcmc.py:377-401
    MetropolisHastings (+ AdaptiveStepsize) geosss/mcmc.py:78-173
    IndependenceSampler                    geosss/mcmc.py:176-179
    MixtureRWMHIndependenceSampler         geosss/mcmc.py:182-234
    SphericalHMC                           geosss/mcmc.py:237-332

Drop-in use is unchanged::

    samples = ShrinkageSphericalSliceSampler(pdf, init_state, seed).sample(n_samples, burnin)

with ``init_state`` of shape ``(d,)`` -> ``(n, d)`` exactly as upstream.  Extension: an
``init_state`` of shape ``(N, d)`` runs N independent chains and returns ``(N, n, d)`` (the
``(chain, draw, dim)`` layout the upstream scripts stack, scripts/protein_reg3d3d.py:346-349).

Same semantics as upstream: the first retained sample is the current state, so
``n_samples + burnin - 1`` transitions run; a float burn-in is a fraction of ``n_samples``;
calling ``sample`` again continues the chain.  Differences, by construction: draws come from a
per-chain Philox4x32-10 stream keyed by ``seed`` and the global chain id (numpy's PCG64 sequence
is not reproduced, so results agree with upstream in distribution, and exactly when both are fed
the same draws -- see ``run_tape``); ``dtype='float32'`` (default) uses the throughput kernels,
``dtype='float64'`` the literal ones.
"""

import ctypes as C

import numpy as np
import torch

from . import _device, _lib


def determine_burnin(n_samples, burnin):
    """Float in [0, 1]: fraction of n_samples; anything else: a count  (mcmc.py:13-19)."""
    if isinstance(burnin, float):
        assert 0 <= burnin <= 1.0
        return int(burnin * n_samples)
    assert burnin >= 0
    return int(burnin)


class PhiloxStream:
    """What ``sampler.rng`` holds here: the key and position of the counter-based stream."""

    def __init__(self, seed, chain_offset):
        self.seed, self.chain_offset, self.steps_done = seed, chain_offset, 0

    def __repr__(self):
        return f"PhiloxStream(seed={self.seed:#x}, chain_offset={self.chain_offset}, steps_done={self.steps_done})"


class Sampler:
    """Base class: state handling and the ``sample`` driver (mcmc.py:22-77)."""

    KIND = None

    def __init__(self, distribution, initial_state, seed=None, *, dtype="float32", path="auto", device=None,
                 chain_offset=0):
        self.target = distribution
        self._torch_io = isinstance(initial_state, torch.Tensor)
        # CUDA initial states stay on the device: no host round trip on the way into the sampler
        self._cuda_state0 = initial_state.detach() if self._torch_io and initial_state.is_cuda else None
        if not self._torch_io:
            initial_state = np.asarray(initial_state)
        if initial_state.ndim not in (1, 2):
            raise ValueError("initial_state must have shape (d,) or (N, d)")
        self._single = initial_state.ndim == 1
        if self._cuda_state0 is not None:
            self._host_state = None  # fetched from the device on demand (``state``)
            self.N, self.d = (1, initial_state.shape[0]) if self._single else tuple(initial_state.shape)
        else:
            x0 = initial_state.detach().numpy() if self._torch_io else np.asarray(initial_state)
            # as upstream, the array is stored, not copied (mcmc.py:43); it is only ever replaced, never written to
            self._host_state = np.asarray(x0, dtype=float).reshape(-1, x0.shape[-1])
            self.N, self.d = self._host_state.shape
        if distribution.d is not None and self.d != distribution.d:
            raise ValueError(f"initial_state has dimension {self.d}, target has {distribution.d}")
        self._code, self._tdtype, self._npdtype = _device.dtype_code(dtype)
        if path not in ("auto", "fast", "literal"):
            raise ValueError("path must be 'auto', 'fast' or 'literal'")
        self._path_req = path
        self._device_req = device if device is not None else (initial_state.device if self._torch_io and initial_state.is_cuda else None)
        self.rng = PhiloxStream(_device.seed_to_u64(seed), int(chain_offset))
        self._h = None
        self._n_reject = 0
        self._rej_seen = 0
        self._steps_counted = 0

    # ---- device sampler ------------------------------------------------------------------
    def _sampler(self):
        if self._h is None:
            self.device = _device.resolve_device(self._device_req)
            lib = _lib.lib()
            tgt = self.target._handle(self.device)
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                want_fast = self._path_req in ("auto", "fast") and self._code == _lib.F32
                rc = _lib.E_UNSUPPORTED
                if want_fast:
                    rc = lib.geosss_sampler_create(tgt, self.KIND, self._code, _lib.PATH_FAST, self.N, self.rng.seed,
                                                   self.rng.chain_offset, C.byref(h))
                    self.path = "fast"
                if rc == _lib.E_UNSUPPORTED and self._path_req != "fast":
                    rc = lib.geosss_sampler_create(tgt, self.KIND, self._code, _lib.PATH_LITERAL, self.N, self.rng.seed,
                                                   self.rng.chain_offset, C.byref(h))
                    self.path = "literal"
                _lib.check(rc)
                self._h = h
                if self._cuda_state0 is not None:
                    x0 = self._cuda_state0.to(device=self.device, dtype=self._tdtype).reshape(self.N, self.d).contiguous()
                    _lib.check(lib.geosss_sampler_set_state(h, x0.data_ptr(), _device.stream_ptr(self.device)))
                    torch.cuda.current_stream(self.device).synchronize()  # x0 may be a temporary
                    self._cuda_state0 = None
                else:
                    _lib.check(lib.geosss_sampler_set_state_host(h, _lib.dptr(np.ascontiguousarray(self._host_state))))
        return self._h

    def _state_device(self):
        """(N, d) CUDA tensor of the current positions in the sampler's dtype (device-to-device copy)."""
        h = self._sampler()
        x = torch.empty((self.N, self.d), dtype=self._tdtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_sampler_get_state(h, x.data_ptr(), None, _device.stream_ptr(self.device)))
        return x

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None:
            try:
                with torch.cuda.device(self.device):
                    _lib.lib().geosss_sampler_destroy(h)
            except Exception:
                pass

    # ---- reference attributes --------------------------------------------------------------
    @property
    def state(self):
        """Current state(s), float64 on the host: (d,) for one chain, (N, d) otherwise."""
        if self._h is None and self._host_state is None:
            self._host_state = self._cuda_state0.detach().cpu().numpy().astype(float).reshape(self.N, self.d)
        if self._h is not None:
            buf = np.empty((self.N, self.d))
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_get_state_host(self._h, _lib.dptr(buf), None))
            self._host_state = buf
        return self._host_state[0] if self._single else self._host_state

    def _positions(self):
        """(N, d) float64 positions (``state`` of the HMC subclass also carries momenta)."""
        return Sampler.state.fget(self).reshape(self.N, self.d)

    @state.setter
    def state(self, value):
        value = np.array(value, dtype=float).reshape(self.N, self.d)
        self._host_state = value
        self._cuda_state0 = None
        if self._h is not None:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_set_state_host(self._h, _lib.dptr(np.ascontiguousarray(value))))

    @property
    def n_reject_per_chain(self):
        if self._h is None:
            return np.zeros(self.N, dtype=np.uint64)
        buf = np.empty(self.N, dtype=np.uint64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_sampler_get_state_host(self._h, None, buf.ctypes.data_as(C.c_void_p)))
        return buf

    @property
    def n_reject(self):
        """Rejected proposals so far, summed over chains (mcmc.py:355,374,401)."""
        self._sync_counters()
        return self._n_reject

    def _reject_total(self):
        """Sum of the per-chain counters, reduced on the device (one scalar crosses to the host)."""
        cnt = torch.empty(self.N, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_sampler_get_state(self._h, None, cnt.data_ptr(), _device.stream_ptr(self.device)))
        return int(cnt.sum().item())

    def _sync_counters(self):
        if self._h is None:
            return
        total = self._reject_total()
        steps = self.rng.steps_done
        # the reference evaluates log_prob 2 + r times per transition (mcmc.py:364-371,389-397)
        calls = 2 * (steps - self._steps_counted) * self.N + (total - self._rej_seen)
        lp = getattr(self.target, "log_prob", None)
        if calls and hasattr(lp, "add_calls"):
            lp.add_calls(calls)
        self._n_reject, self._rej_seen, self._steps_counted = total, total, steps

    # ---- iteration -------------------------------------------------------------------------
    def __iter__(self):
        return self

    def __next__(self):
        """One transition of every chain; returns the new state(s)."""
        self._run_device(1, None)
        return self.state.copy()

    def _run_device(self, n_steps, samples, thin=1, logp=None, nrej=None, chain_major=False, chain_stride=0,
                    trace=None, direction=None):
        h = self._sampler()
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            if trace is not None:
                v = np.ascontiguousarray(direction, dtype=float)
                assert v.shape == (self.d,)
                _lib.check(lib.geosss_sampler_trace_out(h, _lib.dptr(v), trace.data_ptr()))
            try:
                _lib.check(lib.geosss_sampler_run_layout(
                    h, n_steps, thin, _lib.CHAIN_MAJOR if chain_major else _lib.STEP_MAJOR, chain_stride, 0,
                    None if samples is None else samples.data_ptr(),
                    None if logp is None else logp.data_ptr(), None if nrej is None else nrej.data_ptr(),
                    _device.stream_ptr(self.device)))
            finally:
                if trace is not None:
                    _lib.check(lib.geosss_sampler_trace_out(h, None, None))
        self.rng.steps_done += n_steps

    def sample_trace(self, n_samples, direction=None, burnin=0, thin=1, *, out=None):
        """Trace of the scalar test function ``f(x) = x . direction`` (default: the first coordinate) at
        ``n_samples`` successive kept states of every chain, WITHOUT retaining the states: ``(N, n_samples)`` CUDA
        tensor of the sampler's dtype (``(n_samples,)`` for one chain).  This is what the upstream ESS scripts slice out
        of fully retained chains (scripts/ess_mixture_vMF.py:233, scripts/ess_curve_vMF.py:267; ``x . mode`` in
        scripts/bingham.py:50-52); written by the sampler kernels, 4 bytes per kept transition instead of 4 d.
        ``burnin`` transitions run first (a float is a fraction of ``n_samples``), then ``thin`` transitions per kept state."""
        assert n_samples > 0 and thin >= 1
        burnin = determine_burnin(n_samples, burnin)
        self._sampler()
        v = np.zeros(self.d)
        if direction is None:
            v[0] = 1.0
        else:
            v[:] = np.asarray(direction, dtype=float)
        shape = (self.N, n_samples)
        if out is None:
            out = torch.empty(shape, dtype=self._tdtype, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != self._tdtype or not out.is_cuda or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous CUDA {self._tdtype} tensor of shape {shape}")
        if burnin:
            self._run_device(burnin, None)
        self._run_device(n_samples * thin, None, thin=thin, chain_major=True, trace=out, direction=v)
        self._sync_counters()
        return out[0] if self._single else out

    def sample(self, n_samples, burnin=0, return_all_samples=False, *, out=None):
        """Markov chain(s) of the requested length (mcmc.py:55-77).

        Returns float arrays of the sampler's dtype: ``(n, d)`` for one chain, ``(N, n, d)`` for N chains.
        numpy in -> numpy out through pinned host memory (step-major chunks, device->host copies overlapped
        with sampling; the result is the transposed view of that ``(n, N, d)`` buffer); a CUDA tensor as
        ``initial_state`` -> contiguous ``(N, n, d)`` CUDA tensor written chain-major by the kernels, no host copy.
        ``out``: optional preallocated buffer to fill: pinned ``(n, N, d)`` or CUDA ``(N, n, d)``.
        """
        assert n_samples > 0
        burnin = determine_burnin(n_samples, burnin)
        n_out = n_samples + burnin if return_all_samples else n_samples
        keeps_current = return_all_samples or burnin == 0
        n_skip = 0 if keeps_current else burnin - 1
        n_new = n_out - 1 if keeps_current else n_out
        h = self._sampler()
        lib = _lib.lib()
        on_device = self._torch_io and self._device_req is not None
        if out is not None:
            on_device = out.is_cuda
        shape = (self.N, n_out, self.d) if on_device else (n_out, self.N, self.d)
        if out is not None and (tuple(out.shape) != shape or out.dtype != self._tdtype or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous {self._tdtype} tensor of shape {shape}")
        if on_device:
            buf = out if out is not None else torch.empty(shape, dtype=self._tdtype, device=self.device)
            if n_skip:
                self._run_device(n_skip, None)
            # one contiguous (N, n_out, d) block: slot 0 of every chain is its current state when it is kept, its new states follow
            with torch.cuda.device(self.device):
                _lib.check(lib.geosss_sampler_run_block(h, n_new, 1, 1 if keeps_current else 0, buf.data_ptr(), _device.stream_ptr(self.device)))
            self.rng.steps_done += n_new
            res = buf
        else:
            buf = out if out is not None else torch.empty(shape, dtype=self._tdtype, pin_memory=True)
            if keeps_current:
                cur = self._host_state if (self.rng.steps_done == 0 and self._host_state is not None) else self._positions()
                np.copyto(buf[0].numpy(), cur, casting="same_kind")  # one pass, float64 -> sampler dtype
            dst = buf[1:] if keeps_current else buf
            with torch.cuda.device(self.device):
                _lib.check(lib.geosss_sampler_run_host(h, n_skip, n_new, dst.data_ptr() if n_new else None))
            self.rng.steps_done += n_skip + n_new
            res = buf.numpy().transpose(1, 0, 2) if not self._torch_io else buf.permute(1, 0, 2)
        self._sync_counters()
        return res[0] if self._single else res

    # ---- parity mode ---------------------------------------------------------------------
    def run_tape(self, tapes, n_steps):
        """Run ``n_steps`` transitions with the draws taken from ``tapes`` (one flat float64
        array per chain, in the reference's consumption order -- see oracle/tape.py).  Literal
        path only.  Returns (samples (N, n_steps, d), rejections (N, n_steps), status (N,))."""
        if self._path_req == "auto":
            self._path_req = "literal"
        h = self._sampler()
        if self.path != "literal":
            raise ValueError("tape mode needs path='literal'")
        if self._single and not isinstance(tapes, (list, tuple)):
            tapes = [tapes]
        assert len(tapes) == self.N
        offs = np.zeros(self.N + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(t) for t in tapes])
        flat = torch.as_tensor(np.concatenate([np.asarray(t, dtype=float) for t in tapes])).to(self.device)
        offs_d = torch.as_tensor(offs).to(self.device)
        samples = torch.empty((n_steps, self.N, self.d), dtype=self._tdtype, device=self.device)
        nrej = torch.zeros((n_steps, self.N), dtype=torch.int32, device=self.device)
        status = torch.zeros(self.N, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_sampler_run_tape(h, n_steps, flat.data_ptr(), offs_d.data_ptr(), samples.data_ptr(),
                                                          nrej.data_ptr(), status.data_ptr(), _device.stream_ptr(self.device)))
            torch.cuda.synchronize(self.device)
        self.rng.steps_done += n_steps
        return (samples.permute(1, 0, 2).cpu().numpy(), nrej.t().cpu().numpy(), status.cpu().numpy())


class RejectionSphericalSliceSampler(Sampler):
    """GeoSSS-reject: proposals uniform on the whole great circle until one falls in the slice."""

    KIND = _lib.REJECT


class ShrinkageSphericalSliceSampler(RejectionSphericalSliceSampler):
    """GeoSSS-shrink: the proposal bracket shrinks towards the current state on rejections."""

    KIND = _lib.SHRINK


class MetropolisHastings(Sampler):
    """Random-walk Metropolis-Hastings on the sphere with burn-in step-size adaptation
    (mcmc.py:78-173): radius ``sqrt(2 Gamma(d/2))``, isotropic Gaussian perturbation of scale
    ``stepsize``, projection back to the sphere; during the first ``burnin`` transitions of ``sample``
    every chain multiplies its own step size by 1.02 (accepted) or 0.98 (rejected)."""

    KIND = _lib.RWMH

    def __init__(self, distribution, initial_state, seed=None, stepsize=1e-1, **kwargs):
        super().__init__(distribution, initial_state, seed, **kwargs)
        stepsize = float(stepsize)
        assert stepsize > 0.0
        self._cfg = {"stepsize": stepsize, "alpha": -1.0, "n_leap": 0}
        self._adapt = 0
        self._n_accept = 0

    # ---- device sampler ------------------------------------------------------------------
    def _sampler(self):
        fresh = self._h is None
        h = super()._sampler()
        if fresh:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_mh_configure(h, self._cfg["stepsize"], self._cfg["alpha"], self._cfg["n_leap"]))
                _lib.check(_lib.lib().geosss_sampler_mh_reset(h, self._adapt))
        return h

    def _mh_host(self, want):
        """Host copies of per-chain step sizes / random-walk counts / momenta."""
        step = np.empty(self.N) if want == "stepsize" else None
        walk = np.empty(self.N, dtype=np.uint64) if want == "n_walk" else None
        mom = np.empty((self.N, self.d)) if want == "momentum" else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_sampler_mh_get_host(
                self._h, None if step is None else _lib.dptr(step), None if walk is None else walk.ctypes.data_as(C.c_void_p),
                None if mom is None else _lib.dptr(mom)))
        return step if step is not None else (walk if walk is not None else mom)

    # ---- reference attributes --------------------------------------------------------------
    @property
    def stepsize(self):
        """Current step size: a float for one chain, ``(N,)`` otherwise (each chain adapts its own)."""
        if self._h is None:
            return self._cfg["stepsize"] if self._single else np.full(self.N, self._cfg["stepsize"])
        s = self._mh_host("stepsize")
        return float(s[0]) if self._single else s

    @stepsize.setter
    def stepsize(self, value):
        value = float(value)
        assert value > 0.0
        self._cfg["stepsize"] = value
        if self._h is not None:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_mh_configure(self._h, value, -1.0, 0))

    def reset(self, burnin):
        """Restart the adaptation window (AdaptiveStepsize.reset, mcmc.py:101-104)."""
        self._adapt = int(burnin)
        assert self._adapt >= 0
        if self._h is not None:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_mh_reset(self._h, self._adapt))

    @property
    def n_accept_per_chain(self):
        return self.n_reject_per_chain  # the device counter holds accepted proposals for these kinds

    @property
    def n_accept(self):
        """Accepted proposals so far, summed over chains (mcmc.py:133,158)."""
        self._sync_counters()
        return self._n_accept

    @property
    def n_reject(self):
        raise AttributeError("n_reject belongs to the slice samplers; this sampler counts n_accept")

    _GRADS_PER_STEP = 0

    def _sync_counters(self):
        if self._h is None:
            return
        self._n_accept = self._reject_total()
        steps = self.rng.steps_done
        new = (steps - self._steps_counted) * self.N
        # upstream evaluates log_prob twice per transition (mcmc.py:149-153,287-289) and, for HMC,
        # the gradient n_steps + 1 times (mcmc.py:298-311)
        lp, gr = getattr(self.target, "log_prob", None), getattr(self.target, "gradient", None)
        if new and hasattr(lp, "add_calls"):
            lp.add_calls(2 * new)
        if new and self._GRADS_PER_STEP and hasattr(gr, "add_calls"):
            gr.add_calls(self._GRADS_PER_STEP * new)
        self._steps_counted = steps

    def sample(self, n_samples, burnin=0, return_all_samples=False, *, out=None):
        """As ``Sampler.sample``; the burn-in is also the adaptation window (mcmc.py:164-173)."""
        self.reset(determine_burnin(n_samples, burnin))
        return super().sample(n_samples, burnin, return_all_samples, out=out)


class IndependenceSampler(MetropolisHastings):
    """Metropolis-Hastings with uniform proposals on the sphere (mcmc.py:176-179)."""

    KIND = _lib.INDEPENDENCE


class MixtureRWMHIndependenceSampler(MetropolisHastings):
    """With probability ``mixing_probability`` a random-walk proposal, else a uniform one; only the
    random-walk proposals adapt the step size (mcmc.py:182-234).  ``rwmh_stepsize_vals`` (the per-step
    trace upstream appends to) is not kept; ``stepsize`` holds the current values."""

    KIND = _lib.MIXTURE

    def __init__(self, distribution, initial_state, seed=None, stepsize=1e-1, mixing_probability=0.5, **kwargs):
        super().__init__(distribution, initial_state, seed, stepsize=stepsize, **kwargs)
        self.alpha = float(mixing_probability)
        assert 0.0 <= self.alpha <= 1.0
        self._cfg["alpha"] = self.alpha

    @property
    def rwmh_counter(self):
        return 0 if self._h is None else int(self._mh_host("n_walk").sum())

    @property
    def indep_counter(self):
        return self.rng.steps_done * self.N - self.rwmh_counter


class SphericalHMC(MetropolisHastings):
    """Spherical Hamiltonian Monte Carlo (mcmc.py:237-332): tangent Gaussian momenta, ``n_steps``
    geodesic leapfrog steps of size ``stepsize`` driven by ``target.gradient``, Metropolis correction.
    ``state`` is ``[x, v]`` stacked as upstream; ``sample`` returns positions (and momenta on request)."""

    KIND = _lib.HMC

    def __init__(self, distribution, initial_state, seed=None, stepsize=1e-3, n_steps=10, **kwargs):
        super().__init__(distribution, initial_state, seed, stepsize=stepsize, **kwargs)
        self.n_steps = int(n_steps)
        assert self.n_steps > 0
        self._cfg["n_leap"] = self.n_steps
        self._GRADS_PER_STEP = self.n_steps + 1

    @property
    def state(self):
        x = self._positions()
        v = np.zeros_like(x) if self._h is None else self._mh_host("momentum")
        xv = np.hstack([x, v])
        return xv[0] if self._single else xv

    @state.setter
    def state(self, value):
        value = np.array(value, dtype=float).reshape(self.N, -1)
        Sampler.state.fset(self, value[:, :self.d])  # momenta are redrawn at every transition

    def __next__(self):
        self._run_device(1, None)
        return self.state.copy()

    def sample(self, n_samples, burnin=0, return_momenta=False, return_all_samples=False, *, out=None):
        if not return_momenta:
            return super().sample(n_samples, burnin, return_all_samples, out=out)
        assert n_samples > 0
        burnin = determine_burnin(n_samples, burnin)
        self.reset(burnin)
        n_out = n_samples + burnin if return_all_samples else n_samples
        keeps_current = return_all_samples or burnin == 0
        n_skip = 0 if keeps_current else burnin - 1
        n_new = n_out - 1 if keeps_current else n_out
        h = self._sampler()
        pos = torch.empty((self.N, n_out, self.d), dtype=self._tdtype, device=self.device)
        mom = torch.empty_like(pos)
        if keeps_current:
            xv = self.state.reshape(self.N, 2 * self.d)
            pos[:, 0] = torch.as_tensor(xv[:, :self.d], dtype=self._tdtype, device=self.device)
            mom[:, 0] = torch.as_tensor(xv[:, self.d:], dtype=self._tdtype, device=self.device)
        if n_skip:
            self._run_device(n_skip, None)
        if n_new:
            first = 1 if keeps_current else 0
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().geosss_sampler_mh_momenta_out(h, mom[:, first:].data_ptr()))
            try:
                self._run_device(n_new, pos[:, first:], chain_major=True, chain_stride=n_out * self.d)
                torch.cuda.synchronize(self.device)
            finally:
                with torch.cuda.device(self.device):
                    _lib.check(_lib.lib().geosss_sampler_mh_momenta_out(h, None))
        self._sync_counters()
        if not (self._torch_io and self._device_req is not None):
            pos, mom = pos.cpu(), mom.cpu()
            if not self._torch_io:
                pos, mom = pos.numpy(), mom.numpy()
        return (pos[0], mom[0]) if self._single else (pos, mom)
