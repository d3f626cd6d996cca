"""
Diagnostics on retained samples: multi-chain effective sample size and moments.

The upstream scripts compute relative ESS with ``arviz.ess`` over ``(chains, draws)`` traces of
the first coordinate (scripts/ess_mixture_vMF.py:233-268) or of ``x . mode``
(scripts/bingham.py:50-52); the library's own helpers are ``acf`` / ``IAT`` / ``n_eff``
(geosss/utils.py:91-134).  Here the per-chain autocovariances are accumulated on the GPU
(csrc/diag.cuh); only ``max_lag + 2`` doubles per GPU are all-reduced when chains are sharded,
and the Geyer truncation (Vehtari et al. 2021, as restated in oracle/ess.py) runs on the host.
"""

import math

import numpy as np
import torch

from . import _device, _lib


def ess_from_partials(partials, n_draws, n_chains):
    """Finish the estimator from summed partials (max_lag autocovariance sums, sum of chain means,
    sum of squared chain means).  Returns (ess, relative_ess)."""
    partials = np.asarray(partials, dtype=float)
    L = len(partials) - 2
    n, m = float(n_draws), float(n_chains)
    acov = partials[:L] / m
    mean_var = acov[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if n_chains > 1:
        s1, s2 = partials[L], partials[L + 1]
        var_plus += (s2 - s1 * s1 / m) / (m - 1.0)
    rho = 1.0 - (mean_var - acov) / var_plus
    rho[0] = 1.0
    tau, prev, t = -1.0, math.inf, 0
    while t + 1 < L:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        prev = pair
        tau += 2.0 * pair
        t += 2
    tau = max(tau, 1.0 / math.log10(m * n))
    ess = m * n / tau
    return ess, ess / (m * n)


def ess_partials(samples, direction=None, max_lag=None, chain_major=True):
    """Device partial sums.  ``samples``: CUDA tensor, ``(n_chains, n_draws, d)`` as returned by
    ``sample()`` (``chain_major=True``, default) or step-major ``(n_draws, n_chains, d)``.
    Returns a CUDA float64 tensor of length max_lag + 2."""
    if not (isinstance(samples, torch.Tensor) and samples.is_cuda):
        raise ValueError("ess_partials needs a CUDA tensor; there is no CPU fallback")
    samples = samples.contiguous()
    if chain_major:
        N, n, d = samples.shape
    else:
        n, N, d = samples.shape
    code, tdt, _ = _device.dtype_code(samples.dtype)
    max_lag = int(min(n, max_lag if max_lag is not None else n // 2))
    v = torch.zeros(d, dtype=tdt, device=samples.device)
    if direction is None:
        v[0] = 1.0
    else:
        v.copy_(torch.as_tensor(np.asarray(direction, dtype=float)))
    out = torch.empty(max_lag + 2, dtype=torch.float64, device=samples.device)
    with torch.cuda.device(samples.device):
        _lib.check(_lib.lib().geosss_diag_ess_partials(samples.data_ptr(), code,
                                                       _lib.CHAIN_MAJOR if chain_major else _lib.STEP_MAJOR, n, N, d, v.data_ptr(), max_lag,
                                                       out.data_ptr(), _device.stream_ptr(samples.device)))
    return out


def ess(samples, direction=None, max_lag=None, group=None, chain_major=True):
    """Multi-chain ESS of x . direction (default: first coordinate) over ``(n_chains, n_draws, d)``
    samples (or step-major with ``chain_major=False``).  With a torch.distributed
    ``group`` (or the default group when initialised) the partial sums are all-reduced, so every
    rank returns the ESS of ALL chains of all ranks."""
    import torch.distributed as dist
    part = ess_partials(samples, direction, max_lag, chain_major)
    N, n = (samples.shape[0], samples.shape[1]) if chain_major else (samples.shape[1], samples.shape[0])
    total = torch.tensor([float(N)], dtype=torch.float64, device=samples.device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(part, group=group)
        dist.all_reduce(total, group=group)
    return ess_from_partials(part.cpu().numpy(), n, int(total.item()))


def _tau_bulk(acov_sum, mean_sum, mean_sq_sum, n_draws, n_chains):
    """Integrated autocorrelation time from all-reduced partial sums with the truncation rules of
    ``arviz.ess`` (Geyer's initial positive sequence on the pair sums rho(2k) + rho(2k+1), then the initial
    monotone sequence; Vehtari et al. 2021, section 3.2).  Returns (tau, exhausted): ``exhausted`` when the
    lags supplied ended before a pair sum turned non-positive (the caller then asks the device for more)."""
    n, m = float(n_draws), float(n_chains)
    acov = np.asarray(acov_sum, dtype=float) / m
    L = len(acov)
    mean_var = acov[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if n_chains > 1:
        var_plus += (mean_sq_sum - mean_sum * mean_sum / m) / (m - 1.0)
    rho = 1.0 - (mean_var - acov) / var_plus
    rho[0] = 1.0
    # pair k >= 1 is formed while its predecessor is positive and 2k - 1 < n - 3; the last pair formed (index J)
    # contributes only its even lag
    J, last_pair, exhausted = 0, rho[0] + rho[1], False
    while 2 * (J + 1) - 1 < n_draws - 3 and last_pair > 0.0:
        if 2 * (J + 1) + 1 >= L:
            exhausted = True
            break
        J += 1
        last_pair = rho[2 * J] + rho[2 * J + 1]
    if J == 0:
        tau = 0.0
    else:
        full = np.minimum.accumulate(rho[0:2 * J:2] + rho[1:2 * J:2])  # monotone: running minimum of the pair sums
        even = rho[2 * J]
        tau = -1.0 + 2.0 * float(full.sum()) + (even if last_pair >= 0.0 else max(even, 0.0))
    return max(tau, 1.0 / math.log10(m * n)), exhausted


def bulk_ess(trace, group=None, relative=False, local_only=False):
    """Rank-normalised split bulk-ESS of a scalar trace, the estimator the upstream scripts obtain from
    ``arviz.ess`` (default method "bulk"; scripts/bingham.py:28-70, scripts/ess_curve_vMF.py:97-130,
    scripts/vMF_diagnostics.py:451-492), after Vehtari et al. (2021): every chain is split in two halves, the
    draws are replaced by the normal scores of their ranks over ALL draws of all chains, the autocovariances
    of the half chains are accumulated on the GPU (``csrc/diag.cuh``) and the initial-monotone-sequence
    estimator finishes on the host.

    ``trace``: CUDA tensor ``(n_chains, n_draws)`` -- e.g. ``sampler.sample_trace(...)`` or
    ``samples[..., 0]``.  With ``torch.distributed`` initialised the ranks are taken over the chains of all
    processes (all_gather of the scalar traces) and the partial sums are all-reduced, so every process
    returns the ESS of all chains (``local_only`` keeps the call to this process's chains).  Returns ESS, or
    ESS / (chains x draws) with ``relative``."""
    import torch.distributed as dist
    if not (isinstance(trace, torch.Tensor) and trace.is_cuda and trace.dim() == 2):
        raise ValueError("bulk_ess needs a CUDA tensor (n_chains, n_draws); there is no CPU fallback")
    N, n = trace.shape
    half = n // 2
    if half < 4:
        raise ValueError("bulk_ess needs at least 8 draws per chain")
    local = torch.cat([trace[:, :half], trace[:, n - half:]], dim=0).to(torch.float32).contiguous()  # (2N, half)
    distributed = not local_only and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if distributed:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        counts = torch.zeros(world, dtype=torch.int64, device=trace.device)
        counts[rank] = local.shape[0]
        dist.all_reduce(counts, group=group)
        biggest = int(counts.max())
        pad = local if local.shape[0] == biggest else torch.cat([local, local.new_zeros(biggest - local.shape[0], half)])
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        everything = torch.cat([p[:int(c)] for p, c in zip(parts, counts)])
        first = int(counts[:rank].sum())
    else:
        everything, first = local, 0
    # fractional ranks over all draws (ties share the mean of their ranks), then normal scores
    flat = everything.reshape(-1)
    S = flat.numel()
    srt, order = torch.sort(flat)
    new_run = torch.ones(S, dtype=torch.bool, device=flat.device)
    new_run[1:] = srt[1:] != srt[:-1]
    run_first = torch.nonzero(new_run).squeeze(1)
    run_last = torch.cat([run_first[1:], run_first.new_tensor([S])]) - 1
    run_id = torch.cumsum(new_run, 0) - 1
    mean_rank = 0.5 * (run_first + run_last).to(torch.float64)[run_id] + 1.0
    ranks = torch.empty(S, dtype=torch.float64, device=flat.device)
    ranks[order] = mean_rank
    z = torch.special.ndtri((ranks - 0.375) / (S + 0.25)).to(torch.float32).reshape(everything.shape)
    mine = z[first:first + local.shape[0]].contiguous().unsqueeze(-1)  # (2N, half, 1): this process's half chains
    m_total = everything.shape[0]
    lags = min(half, 128)
    while True:
        part = ess_partials(mine, np.ones(1), lags, chain_major=True)
        if distributed:
            dist.all_reduce(part, group=group)
        p = part.cpu().numpy()
        tau, exhausted = _tau_bulk(p[:lags], p[lags], p[lags + 1], half, m_total)
        if not exhausted or lags >= half:
            break
        lags = min(half, 2 * lags)
    return 1.0 / tau if relative else m_total * half / tau


class StreamingESS:
    """Multi-chain autocovariance ESS of a scalar trace that is never retained: chunks of the trace (as the sampler
    kernels write them, ``Sampler.sample_trace``) are folded into per-chain accumulators on the GPU
    (``csrc/diag.cuh:k_ess_stream_update``: running sum, lagged products for lags < ``max_lag``, first and last values),
    so a run of any length costs ``n_chains x (1 + 3 max_lag)`` doubles.  ``result()`` equals ``ess()`` of the whole
    retained trace with the same ``max_lag`` (tested to 1e-9 against the float64 oracle); with torch.distributed the
    partial sums are all-reduced, so every rank gets the ESS of all chains."""

    def __init__(self, n_chains, max_lag, device=None):
        self.device = _device.resolve_device(device)
        self.n_chains, self.max_lag, self.n = int(n_chains), int(max_lag), 0
        self.acc = torch.zeros((self.n_chains, 1 + 3 * self.max_lag), dtype=torch.float64, device=self.device)

    def update(self, trace):
        """``trace``: CUDA tensor ``(n_chains, n_new)``, contiguous."""
        if not (isinstance(trace, torch.Tensor) and trace.is_cuda and trace.dim() == 2 and trace.shape[0] == self.n_chains):
            raise ValueError("update needs a CUDA tensor (n_chains, n_new)")
        trace = trace.contiguous()
        code, _, _ = _device.dtype_code(trace.dtype)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_diag_stream_update(trace.data_ptr(), code, _lib.CHAIN_MAJOR, trace.shape[1], self.n_chains, self.n,
                                                            self.max_lag, self.acc.data_ptr(), _device.stream_ptr(self.device)))
        self.n += int(trace.shape[1])
        return self

    def partials(self):
        out = torch.empty(self.max_lag + 2, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().geosss_diag_stream_finish(self.acc.data_ptr(), self.n_chains, self.n, self.max_lag, out.data_ptr(),
                                                            _device.stream_ptr(self.device)))
        return out

    def result(self, group=None):
        """(ESS, relative ESS) over all chains (of all ranks)."""
        import torch.distributed as dist
        part = self.partials()
        total = torch.tensor([float(self.n_chains)], dtype=torch.float64, device=self.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(part, group=group)
            dist.all_reduce(total, group=group)
        return ess_from_partials(part.cpu().numpy(), self.n, int(total.item()))


def streaming_ess(sampler, n_samples, direction=None, burnin=0, chunk=500, max_lag=250, group=None):
    """ESS of ``x . direction`` over ``n_samples`` transitions of every chain of ``sampler`` without retaining states or
    trace: the run is cut into chunks of ``chunk`` transitions whose scalar trace (4 bytes per transition and chain) is
    folded into a ``StreamingESS`` and overwritten by the next chunk.  Returns (ESS, relative ESS)."""
    n_samples, chunk = int(n_samples), int(min(chunk, n_samples))
    if burnin:
        sampler._run_device(determine_burnin_count(n_samples, burnin), None)
    sampler._sampler()
    acc = StreamingESS(sampler.N, min(max_lag, n_samples), sampler.device)
    buf = None
    done = 0
    while done < n_samples:
        n = min(chunk, n_samples - done)
        if buf is None or buf.shape[1] != n:
            buf = torch.empty((sampler.N, n), dtype=sampler._tdtype, device=acc.device)
        tr = sampler.sample_trace(n, direction, out=buf)
        acc.update(tr.reshape(sampler.N, n))
        done += n
    return acc.result(group)


def determine_burnin_count(n_samples, burnin):
    from .mcmc import determine_burnin
    return determine_burnin(n_samples, burnin)


def path_stats(samples, modes=None, sign_labels=False, group=None, chain_major=True):
    """One pass over retained draws on the GPU: label populations, hops and geodesic steps.

    ``samples``: CUDA tensor ``(n_chains, n_draws, d)`` (or step-major with ``chain_major=False``).
    ``modes``: ``(K, d)`` directions; a draw's label is ``argmax_k x . mode_k``
    (scripts/vMF_diagnostics.py:333-336) or, with ``sign_labels`` and one direction, the sign of
    ``x . mode`` (scripts/bingham.py:23-25).  Returns a dict with ``counts`` (per label; sign labels
    are ordered -, 0, +), ``hops``, ``pairs`` (successive draws compared), ``distance_sum``; the sums
    are all-reduced over ``group`` when torch.distributed is initialised."""
    import torch.distributed as dist
    if not (isinstance(samples, torch.Tensor) and samples.is_cuda):
        raise ValueError("path_stats needs a CUDA tensor; there is no CPU fallback")
    samples = samples.contiguous()
    if chain_major:
        N, n, d = samples.shape
    else:
        n, N, d = samples.shape
    code, _, _ = _device.dtype_code(samples.dtype)
    K = 0
    m = None
    if modes is not None:
        m = torch.as_tensor(np.ascontiguousarray(np.atleast_2d(np.asarray(modes, dtype=float)))).to(samples.device)
        K = m.shape[0]
        if m.shape[1] != d:
            raise ValueError("modes must have shape (K, d)")
    if sign_labels and K != 1:
        raise ValueError("sign labels need exactly one direction")
    L = 3 if sign_labels else K
    out = torch.empty(L + 3, dtype=torch.float64, device=samples.device)
    with torch.cuda.device(samples.device):
        _lib.check(_lib.lib().geosss_diag_path_stats(samples.data_ptr(), code, _lib.CHAIN_MAJOR if chain_major else _lib.STEP_MAJOR,
                                                     n, N, d, m.data_ptr() if K else None, K, int(bool(sign_labels)),
                                                     out.data_ptr(), _device.stream_ptr(samples.device)))
    out[L + 2] = float(N) * (n - 1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(out, group=group)
    o = out.cpu().numpy()
    return {"counts": o[:L].copy(), "hops": o[L], "pairs": o[L + 2], "distance_sum": o[L + 1]}


def mode_kl(samples, pdf, **kwargs):
    """KL divergence between the sampled mode populations and the mixture weights
    (scripts/vMF_diagnostics.py:331-342): populations from ``argmax_k x . mu_k``, empty modes at 1e-100."""
    modes = np.array([p.mu for p in pdf.pdfs])
    c = path_stats(samples, modes, **kwargs)["counts"]
    p = np.full(len(modes), 1e-100)
    p[c > 0] = c[c > 0]
    p[c > 0] /= p.sum()
    return float(p @ np.log(p / np.asarray(pdf.weights)))


def hopping_frequency(samples, pdf, **kwargs):
    """Fraction of successive draws on different sides of the Bingham mode
    (scripts/bingham.py:23-25: ``mean(diff(sign(x . mode)) != 0)``), over all chains."""
    s = path_stats(samples, np.asarray(pdf.mode, dtype=float)[None], sign_labels=True, **kwargs)
    return float(s["hops"] / s["pairs"])


def mean_geodesic_step(samples, **kwargs):
    """Average great-circle distance of successive draws (scripts/curve_vMF.py:379-389)."""
    s = path_stats(samples, **kwargs)
    return float(s["distance_sum"] / s["pairs"])
