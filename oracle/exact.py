"""
Exact i.i.d. samplers used as statistical references (oracle; test infrastructure only).

``wood_vmf`` restates Wood's (1994) rejection sampler for the von Mises-Fisher distribution,
the algorithm behind geosss/rand.py:31-66 (``sample_vMF``), vectorised over draws and with a
seeded Generator instead of the global numpy state.  The tangent component is placed around
the mode with a Householder reflection instead of upstream's SVD rotation; the distribution
is the same because the tangent direction is uniform.
"""

import numpy as np


def wood_vmf(mu, size, seed=None):
    """size i.i.d. draws from vMF(mu) on S^{d-1} (density ~ exp(mu . x))."""
    gen = np.random.default_rng(seed)
    mu = np.asarray(mu, dtype=float)
    d = mu.size
    p = d - 1
    kappa = np.linalg.norm(mu)
    b0 = (-2 * kappa + np.sqrt(4 * kappa**2 + p**2)) / p
    x0 = (1 - b0) / (1 + b0)
    c = kappa * x0 + p * np.log(1 - x0**2)
    w = np.empty(size)
    todo = np.arange(size)
    while todo.size:
        z = gen.beta(0.5 * p, 0.5 * p, todo.size)
        u = np.log(gen.random(todo.size))
        cand = (1 - (1 + b0) * z) / (1 - (1 - b0) * z)
        ok = kappa * cand + p * np.log(1 - x0 * cand) - c >= u
        w[todo[ok]] = cand[ok]
        todo = todo[~ok]
    v = gen.standard_normal((size, p))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    x = np.hstack([np.sqrt(1 - w**2)[:, None] * v, w[:, None]])  # around the north pole e_d
    # Householder reflection taking e_d to mu/kappa
    e = np.zeros(d)
    e[-1] = 1.0
    h = e - mu / kappa
    hn = h @ h
    if hn > 1e-30:
        x = x - 2 * np.outer(x @ h, h) / hn
    return x


def wood_vmf_mixture(mus, weights, size, seed=None):
    """size i.i.d. draws from a vMF mixture."""
    gen = np.random.default_rng(seed)
    mus = np.asarray(mus, dtype=float)
    w = np.asarray(weights, dtype=float)
    comp = gen.choice(len(mus), size=size, p=w / w.sum())
    out = np.empty((size, mus.shape[1]))
    for k in range(len(mus)):
        idx = np.nonzero(comp == k)[0]
        if idx.size:
            out[idx] = wood_vmf(mus[k], idx.size, seed=gen.integers(2**63))
    return out


def kent_envelope(A):
    """Parameters of the angular-central-Gaussian envelope of Kent, Ganeiber & Mardia for
    Bingham(A) ~ exp(x^T A x) (geosss/rand.py:132-158): eigenvalues of -A shifted to start at 0,
    eigenvectors, the root b of sum_j 1 / (b + 2 v_j) = 1 and the log of the bound M."""
    from scipy.optimize import brentq
    A = np.asarray(A, dtype=float)
    if A.ndim == 1:
        v, U = -A, None
    else:
        v, U = np.linalg.eigh(-A)
    v = v - np.min(v)
    d = len(v)
    b = float(d) if np.allclose(v, 0.0) else brentq(lambda t: 1 - np.sum(1 / (t + 2 * v)), 1, d)
    log_m = -(d - b) / 2 + (d / 2) * np.log(d / b)
    return v, U, b, log_m


def kent_bingham(A, size, seed=None, return_efficiency=False):
    """size i.i.d. draws from Bingham(A) by ACG rejection (geosss/rand.py:145-193), seeded."""
    gen = np.random.default_rng(seed)
    v, U, b, log_m = kent_envelope(A)
    d = len(v)
    out, tried = [], 0
    need = size
    while need > 0:
        x = gen.standard_normal((need, d)) / np.sqrt(1 + 2 * v / b)
        x /= np.linalg.norm(x, axis=-1)[:, None]
        u = np.sum(np.square(x) * v, axis=-1)
        ok = np.log(gen.random(need)) < -u + (d / 2) * np.log(1 + 2 * u / b) - log_m
        out.append(x[ok])
        tried += need
        need -= int(ok.sum())
    x = np.concatenate(out, 0)[:size]
    if U is not None:
        x = x @ U.T
    return (x, size / tried) if return_efficiency else x
