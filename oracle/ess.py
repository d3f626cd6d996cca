"""
Effective-sample-size estimators (oracle; test infrastructure only).

``acf_fft`` / ``iat`` / ``n_eff`` restate geosss/utils.py:109-134 (the library's own
diagnostics).  ``ess_multichain`` restates the multi-chain autocorrelation ESS of
Vehtari et al. (2021, "Rank-normalization, folding, and localization", eqs. 10-13:
split-free variant, Geyer initial-positive + initial-monotone truncation) that
the upstream scripts obtain from ``arviz.ess`` (scripts/vMF_diagnostics.py:451-492).
arviz is not installed here, so THAT estimator is "parity unpinned"; the device
implementation is checked against this restatement only.
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
