"""
Effective-sample-size estimators (oracle; test infrastructure only).

``acf_fft`` / ``iat`` / ``n_eff`` restate geosss/utils.py:109-134 (the library's own
diagnostics).  ``ess_multichain`` restates the multi-chain autocorrelation ESS of
Vehtari et al. (2021, "Rank-normalization, folding, and localization", eqs. 10-13:
split-free variant, Geyer initial-positive + initial-monotone truncation) that
the upstream scripts obtain from ``arviz.ess`` (scripts/vMF_diagnostics.py:451-492).
arviz is not installed here, so THAT estimator is "parity unpinned"; the device
implementation is checked against this restatement only.

``bulk_ess`` restates what the upstream scripts actually call -- ``arviz.ess(..., method="bulk")``
(the default; scripts/bingham.py:28-70, scripts/ess_curve_vMF.py:97-130, scripts/vMF_diagnostics.py:451-492;
arviz 0.20.0 pinned in requirements.txt:1): split every chain in two halves, replace the values by the
normal scores of their fractional ranks over ALL draws ((r - 3/8) / (S + 1/4), average ranks for ties),
FFT autocovariance per half chain, Geyer's initial positive / initial monotone sequence exactly as arviz's
``_ess`` implements them (Vehtari, Gelman, Simpson, Carpenter, Buerkner 2021, sections 3.2 and 4).  Written
from the published algorithm, independently of the device code (``geosss_b200/diagnostics.py``); pinned by
known answers (i.i.d. draws, AR(1) chains: ESS/S = (1 - phi) / (1 + phi)) and by the relative ESS the
reference publishes for Bingham d=10 (scripts/Bingham.ipynb:254: reject 99.73 %, shrink 15.2 %).
"""

import numpy as np


def acf_fft(x):
    """Autocorrelation by the convolution theorem, circular (utils.py:109-112)."""
    x = np.asarray(x, dtype=float)
    X = np.fft.rfft((x - x.mean()) / x.std())
    return np.real(np.fft.irfft(X.conj() * X))[: len(x) // 2] / len(x)


def iat(x, n=None):
    """Integrated autocorrelation time, Gelman-style pair sums (utils.py:115-129)."""
    ac = acf_fft(x)
    if n:
        ac = ac[:n]
    body = ac[2:-1] if len(ac) % 2 != 0 else ac[2:]
    pair = body.reshape(-1, 2).sum(1)
    neg = np.nonzero(pair < 0)[0]
    T = neg[0]  # raises IndexError when no pair sum is negative, as upstream does
    L = 1 + 2 * T
    return 1.0 + max(2 * np.sum(ac[1: L + 1]), 0.0)


def n_eff(x, n=None):
    return len(x) / iat(x, n)


def autocov_multichain(x, max_lag):
    """Biased (1/n) per-chain autocovariance at lags 0..max_lag-1. x: (chains, draws)."""
    x = np.asarray(x, dtype=float)
    m, n = x.shape
    xc = x - x.mean(axis=1, keepdims=True)
    out = np.empty((m, max_lag))
    for t in range(max_lag):
        out[:, t] = np.sum(xc[:, : n - t] * xc[:, t:], axis=1) / n
    return out


def ess_multichain(x, max_lag=None):
    """ESS of a scalar trace x (chains, draws) pooled over chains."""
    x = np.asarray(x, dtype=float)
    m, n = x.shape
    max_lag = n if max_lag is None else min(max_lag, n)
    acov = autocov_multichain(x, max_lag)
    chain_mean = x.mean(axis=1)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += chain_mean.var(ddof=1)
    rho = 1.0 - (mean_var - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    # Geyer: sum pairs P_t = rho_2t + rho_2t+1 while positive, made monotone
    tau = -1.0
    prev = np.inf
    t = 0
    while t + 1 < max_lag:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        prev = pair
        tau += 2.0 * pair
        t += 2
    tau = max(tau, 1.0 / np.log10(m * n))
    return m * n / tau


# ------------------------------------------------------------------ arviz-style bulk ESS
def split_chains(x):
    """(chains, draws) -> (2 chains, draws // 2): first and last halves (the middle draw of an odd chain is dropped)."""
    x = np.asarray(x, dtype=float)
    half = x.shape[1] // 2
    return np.concatenate([x[:, :half], x[:, x.shape[1] - half:]], axis=0)


def average_ranks(v):
    """1-based ranks of a flat array, ties get the mean of the ranks they span."""
    v = np.asarray(v).ravel()
    order = np.argsort(v, kind="stable")
    sv = v[order]
    # boundaries of runs of equal values
    start = np.r_[True, sv[1:] != sv[:-1]]
    first = np.flatnonzero(start)
    last = np.r_[first[1:], len(v)] - 1
    run = np.cumsum(start) - 1
    mean_rank = 0.5 * (first + last)[run] + 1.0
    out = np.empty(len(v))
    out[order] = mean_rank
    return out


def z_scale(x):
    """Rank-normalise: normal scores of the fractional ranks over all draws (Blom offsets 3/8, 1/4)."""
    from scipy.special import ndtri
    x = np.asarray(x, dtype=float)
    r = average_ranks(x)
    return ndtri((r - 0.375) / (x.size + 0.25)).reshape(x.shape)


def autocov_fft_chains(x):
    """Biased (1/n) autocovariance of every chain at lags 0..n-1 by zero-padded FFT. x: (chains, draws)."""
    x = np.asarray(x, dtype=float)
    n = x.shape[1]
    xc = x - x.mean(axis=1, keepdims=True)
    size = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, size, axis=1)
    return np.fft.irfft(f * np.conj(f), size, axis=1)[:, :n] / n


def geyer_tau(mean_acov, mean_var, var_plus, n_draw, n_total):
    """Integrated autocorrelation time from chain-averaged autocovariances, arviz's truncation rules.
    mean_acov: lags 0..L-1 (L may be < n_draw when the caller knows the sequence ends earlier)."""
    L = len(mean_acov)
    rho = np.zeros(L + 3)
    rho_at = lambda t: 1.0 - (mean_var - mean_acov[t]) / var_plus
    rho_even, rho_odd = 1.0, rho_at(1)
    rho[0], rho[1] = rho_even, rho_odd
    t = 1
    while t < n_draw - 3 and t + 2 < L and rho_even + rho_odd > 0.0:
        rho_even, rho_odd = rho_at(t + 1), rho_at(t + 2)
        if rho_even + rho_odd >= 0.0:
            rho[t + 1], rho[t + 2] = rho_even, rho_odd
        t += 2
    exhausted = rho_even + rho_odd > 0.0 and t + 2 >= L and t < n_draw - 3
    max_t = t - 2
    if rho_even > 0.0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:  # initial monotone sequence
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = 0.5 * (rho[t - 1] + rho[t])
            rho[t + 2] = rho[t + 1]
        t += 2
    tau = -1.0 + 2.0 * np.sum(rho[: max_t + 1]) + np.sum(rho[max_t + 1: max_t + 2])
    return max(tau, 1.0 / np.log10(n_total)), exhausted


def ess_arviz(x):
    """arviz ``_ess`` on (chains, draws): ESS pooled over chains."""
    x = np.asarray(x, dtype=float)
    m, n = x.shape
    acov = autocov_fft_chains(x)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    tau, _ = geyer_tau(acov.mean(axis=0), mean_var, var_plus, n, m * n)
    return m * n / tau


def bulk_ess(x, relative=False):
    """``arviz.ess(method="bulk")`` of a scalar trace x (chains, draws)."""
    z = z_scale(split_chains(x))
    e = ess_arviz(z)
    return e / z.size if relative else e
